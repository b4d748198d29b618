"""A small InfiniteOpt/JuMP-like modelling layer — just enough of the data model that
``src/transform.jl`` walks (infinite parameters and their supports, finite / infinite / semi-infinite /
point variables, derivatives, parameter functions, measures, JuMP's affine / quadratic / nonlinear
expression containers) to express the reference's test and benchmark models in Python.

InfiniteOpt.jl and JuMP are third-party Julia packages that are not in this container; the expression
canonicalisation below follows JuMP's documented rules (affine + affine = affine with ordered terms,
affine × affine = quadratic with ordered terms, anything else = ``GenericNonlinearExpr(head, args)``).
"""
from __future__ import annotations

import numbers
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np


# ---- derivative methods --------------------------------------------------------------------------
@dataclass
class FiniteDifference:
    """InfiniteOpt's default ``FiniteDifference(Backward())``."""
    kind: str = "backward"


@dataclass
class OrthogonalCollocation:
    """``OrthogonalCollocation(n)``: n nodes per interval including both ends (Lobatto) ->
    n-2 internal (generative) supports per interval (transform.jl:22, :579-584)."""
    num_nodes: int = 3


# ---- references ------------------------------------------------------------------------------------
class Ref:
    """``InfiniteOpt.GeneralVariableRef``: identity-hashed handle with JuMP operator overloading."""
    index_type = "?"
    _count = 0

    def __init__(self, model):
        self.model = model
        Ref._count += 1
        self._id = Ref._count

    def __hash__(self): return self._id
    def __eq__(self, o): return self is o

    # arithmetic -> JuMP containers
    def __add__(self, o): return _add(self, o)
    def __radd__(self, o): return _add(o, self)
    def __sub__(self, o): return _sub(self, o)
    def __rsub__(self, o): return _sub(o, self)
    def __mul__(self, o): return _mul(self, o)
    def __rmul__(self, o): return _mul(o, self)
    def __truediv__(self, o): return _div(self, o)
    def __rtruediv__(self, o): return _div(o, self)
    def __pow__(self, p): return _pow(self, p)
    def __neg__(self): return _neg(self)
    def __pos__(self): return self


class InfiniteParameter(Ref):
    def __init__(self, model, group, pos, dependent, lb=None, ub=None, derivative_method=None):
        super().__init__(model)
        self.group, self.pos, self.dependent = group, pos, dependent      # group: 1-based
        self.lb, self.ub = lb, ub
        self.derivative_method = derivative_method or FiniteDifference()
        self.index_type = "DependentParameter" if dependent else "IndependentParameter"

    @property
    def groups(self): return (self.group,)


class FiniteParameter(Ref):
    index_type = "FiniteParameter"

    def __init__(self, model, value):
        super().__init__(model)
        self.value = float(value)

    groups = ()


class ParameterFunction(Ref):
    index_type = "ParameterFunction"

    def __init__(self, model, func, prefs):
        super().__init__(model)
        self.func, self.prefs = func, tuple(prefs)

    @property
    def groups(self): return _groups_of_prefs(self.prefs)


@dataclass
class VarInfo:
    lb: object = None
    ub: object = None
    fix: object = None
    start: object = None
    binary: bool = False
    integer: bool = False


class FiniteVariable(Ref):
    index_type = "FiniteVariable"
    groups = ()

    def __init__(self, model, info: VarInfo):
        super().__init__(model)
        self.info = info


class InfiniteVariable(Ref):
    index_type = "InfiniteVariable"

    def __init__(self, model, prefs, info: VarInfo):
        super().__init__(model)
        self.prefs, self.info = tuple(prefs), info

    @property
    def groups(self): return _groups_of_prefs(self.prefs)

    def __call__(self, *vals):
        """``y(0, x)``: numbers fix a parameter (point / semi-infinite variable), a parameter ref keeps it"""
        assert len(vals) == len(self.prefs)
        fixed = {i: float(v) for i, v in enumerate(vals) if isinstance(v, numbers.Real)}
        if not fixed:
            return self
        if len(fixed) == len(vals):
            return self.model._point(self, tuple(float(v) for v in vals))
        return self.model._semi(self, fixed)


class Derivative(InfiniteVariable):
    index_type = "Derivative"

    def __init__(self, model, arg, pref):
        super().__init__(model, arg.prefs, VarInfo())
        self.arg, self.pref = arg, pref


class SemiInfiniteVariable(Ref):
    index_type = "SemiInfiniteVariable"

    def __init__(self, model, base, fixed: Dict[int, float]):
        super().__init__(model)
        self.base, self.fixed = base, dict(fixed)
        self.info = VarInfo()

    @property
    def prefs(self): return tuple(p for i, p in enumerate(self.base.prefs) if i not in self.fixed)

    @property
    def groups(self): return _groups_of_prefs(self.prefs)


class PointVariable(Ref):
    index_type = "PointVariable"
    groups = ()

    def __init__(self, model, base, values):
        super().__init__(model)
        self.base, self.values = base, tuple(values)
        self.info = VarInfo()


class Measure(Ref):
    index_type = "Measure"

    def __init__(self, model, expr, prefs, supports, coeffs):
        super().__init__(model)
        self.expr = expr
        self.prefs = tuple(prefs)             # one independent parameter, or all parameters of a dependent group
        self.supports = np.asarray(supports)  # (K,) or (n_params, K)
        self.coeffs = np.asarray(coeffs, dtype=np.float64)

    @property
    def group(self): return self.prefs[0].group

    @property
    def groups(self):
        return tuple(g for g in expression_groups(self.expr) if g != self.group)


def _groups_of_prefs(prefs):
    out = []
    for p in prefs:
        if p.group not in out:
            out.append(p.group)
    return tuple(out)


# ---- JuMP expression containers -------------------------------------------------------------------
class AffExpr:
    def __init__(self, terms=None, constant=0.0):
        self.terms: "OrderedDict[Ref, float]" = OrderedDict(terms or {})
        self.constant = float(constant)

    def copy(self): return AffExpr(self.terms, self.constant)
    __add__ = lambda s, o: _add(s, o)
    __radd__ = lambda s, o: _add(o, s)
    __sub__ = lambda s, o: _sub(s, o)
    __rsub__ = lambda s, o: _sub(o, s)
    __mul__ = lambda s, o: _mul(s, o)
    __rmul__ = lambda s, o: _mul(o, s)
    __truediv__ = lambda s, o: _div(s, o)
    __rtruediv__ = lambda s, o: _div(o, s)
    __pow__ = lambda s, p: _pow(s, p)
    __neg__ = lambda s: _neg(s)


class QuadExpr:
    def __init__(self, terms=None, aff=None):
        self.terms: "OrderedDict[Tuple[Ref, Ref], float]" = OrderedDict(terms or {})
        self.aff = aff if aff is not None else AffExpr()

    def copy(self): return QuadExpr(self.terms, self.aff.copy())
    __add__ = lambda s, o: _add(s, o)
    __radd__ = lambda s, o: _add(o, s)
    __sub__ = lambda s, o: _sub(s, o)
    __rsub__ = lambda s, o: _sub(o, s)
    __mul__ = lambda s, o: _mul(s, o)
    __rmul__ = lambda s, o: _mul(o, s)
    __truediv__ = lambda s, o: _div(s, o)
    __rtruediv__ = lambda s, o: _div(o, s)
    __pow__ = lambda s, p: _pow(s, p)
    __neg__ = lambda s: _neg(s)


class NLExpr:
    """``JuMP.GenericNonlinearExpr``: head symbol + argument list"""

    def __init__(self, head: str, args):
        self.head, self.args = head, list(args)

    __add__ = lambda s, o: NLExpr("+", [s, o])
    __radd__ = lambda s, o: NLExpr("+", [o, s])
    __sub__ = lambda s, o: NLExpr("-", [s, o])
    __rsub__ = lambda s, o: NLExpr("-", [o, s])
    __mul__ = lambda s, o: NLExpr("*", [s, o])
    __rmul__ = lambda s, o: NLExpr("*", [o, s])
    __truediv__ = lambda s, o: NLExpr("/", [s, o])
    __rtruediv__ = lambda s, o: NLExpr("/", [o, s])
    __pow__ = lambda s, p: NLExpr("^", [s, p])
    __neg__ = lambda s: NLExpr("-", [s])


def _is_num(v): return isinstance(v, (numbers.Real, np.floating, np.integer))


def _aff(v) -> AffExpr:
    if isinstance(v, AffExpr): return v
    if isinstance(v, Ref): return AffExpr({v: 1.0})
    if _is_num(v): return AffExpr(constant=float(v))
    raise TypeError(type(v))


def _quad(v) -> QuadExpr:
    if isinstance(v, QuadExpr): return v
    return QuadExpr(aff=_aff(v).copy())


def _add_to(terms, key, c):
    terms[key] = terms.get(key, 0.0) + c


def _add(a, b):
    if isinstance(a, NLExpr) or isinstance(b, NLExpr):
        return NLExpr("+", [a, b])
    if isinstance(a, QuadExpr) or isinstance(b, QuadExpr):
        out = _quad(a).copy() if isinstance(a, QuadExpr) else _quad(a)
        qb = _quad(b)
        for k, c in qb.terms.items(): _add_to(out.terms, k, c)
        for k, c in qb.aff.terms.items(): _add_to(out.aff.terms, k, c)
        out.aff.constant += qb.aff.constant
        return out
    out = _aff(a).copy()
    ab = _aff(b)
    for k, c in ab.terms.items(): _add_to(out.terms, k, c)
    out.constant += ab.constant
    return out


def _sub(a, b):
    if isinstance(a, NLExpr) or isinstance(b, NLExpr):
        return NLExpr("-", [a, b])
    return _add(a, _neg(b))


def _neg(a):
    if _is_num(a): return -a
    if isinstance(a, NLExpr): return NLExpr("-", [a])
    return _mul(-1.0, a)


def _pair(u, v, terms):
    """JuMP.UnorderedPair: (u,v) and (v,u) are the same key; the first insertion fixes the order"""
    return (v, u) if (v, u) in terms and (u, v) not in terms else (u, v)


def _mul(a, b):
    if _is_num(a) and _is_num(b): return a * b
    if isinstance(a, NLExpr) or isinstance(b, NLExpr): return NLExpr("*", [a, b])
    if _is_num(a) or _is_num(b):
        c, e = (float(a), b) if _is_num(a) else (float(b), a)
        if isinstance(e, QuadExpr):
            return QuadExpr(OrderedDict((k, c * v) for k, v in e.terms.items()),
                            AffExpr(OrderedDict((k, c * v) for k, v in e.aff.terms.items()), c * e.aff.constant))
        e = _aff(e)
        return AffExpr(OrderedDict((k, c * v) for k, v in e.terms.items()), c * e.constant)
    if isinstance(a, QuadExpr) or isinstance(b, QuadExpr): return NLExpr("*", [a, b])
    la, lb = _aff(a), _aff(b)
    out = QuadExpr()
    for u, cu in la.terms.items():                       # lhs terms outer, rhs terms inner (JuMP's loop order)
        for v, cv in lb.terms.items():
            _add_to(out.terms, _pair(u, v, out.terms), cu * cv)
    for u, cu in la.terms.items():
        if lb.constant: _add_to(out.aff.terms, u, cu * lb.constant)
    for v, cv in lb.terms.items():
        if la.constant: _add_to(out.aff.terms, v, cv * la.constant)
    out.aff.constant = la.constant * lb.constant
    return out


def _div(a, b):
    if _is_num(b): return _mul(a, 1.0 / b)
    return NLExpr("/", [a, b])


def _pow(a, p):
    if _is_num(p) and p == 2 and not isinstance(a, (QuadExpr, NLExpr)): return _mul(a, a)
    if _is_num(p) and p == 1: return a
    return NLExpr("^", [a, p])


def nl(head: str):
    return lambda *args: NLExpr(head, list(args))


sin, cos, tan, exp, log, sqrt, tanh, sinh, cosh, atan, asin, acos = (nl(h) for h in (
    "sin", "cos", "tan", "exp", "log", "sqrt", "tanh", "sinh", "cosh", "atan", "asin", "acos"))


def all_expression_variables(expr) -> List[Ref]:
    out: List[Ref] = []

    def visit(e):
        if isinstance(e, Ref):
            if e not in out: out.append(e)
        elif isinstance(e, AffExpr):
            for k in e.terms: visit(k)
        elif isinstance(e, QuadExpr):
            for (u, v) in e.terms: visit(u); visit(v)
            visit(e.aff)
        elif isinstance(e, NLExpr):
            for a in e.args: visit(a)
    visit(expr)
    return out


def expression_groups(expr) -> Tuple[int, ...]:
    """``parameter_group_int_indices``: sorted union of the groups an expression depends on"""
    gs = set()
    for v in all_expression_variables(expr):
        gs.update(v.groups)
    return tuple(sorted(gs))


def map_expression(f: Callable, expr):
    """``InfiniteOpt.map_expression``: rebuild the expression with every variable v replaced by f(v)"""
    if isinstance(expr, Ref): return f(expr)
    if _is_num(expr): return expr
    if isinstance(expr, AffExpr):
        out = expr.constant
        for v, c in expr.terms.items(): out = out + c * f(v)
        return out
    if isinstance(expr, QuadExpr):
        out = map_expression(f, expr.aff)
        for (u, v), c in expr.terms.items(): out = out + c * f(u) * f(v)
        return out
    return NLExpr(expr.head, [map_expression(f, a) for a in expr.args])


# ---- the model -------------------------------------------------------------------------------------
@dataclass
class ConstraintData:
    expr: object
    lb: float
    ub: float
    restriction: Optional[Callable] = None          # (dict pref -> value) -> bool
    restriction_prefs: Tuple = ()


class InfiniteModel:
    def __init__(self):
        self.param_groups: List[List[InfiniteParameter]] = []   # one list per group
        self.supports: List[np.ndarray] = []                    # per group: (K,) or (n, K)
        self.public: List[np.ndarray] = []                      # independent parameters: public supports
        self.finite_params: List[FiniteParameter] = []
        self.param_funcs: List[ParameterFunction] = []
        self.finite_vars: List[FiniteVariable] = []
        self.infinite_vars: List[InfiniteVariable] = []
        self.derivatives: List[Derivative] = []
        self.semi_vars: List[SemiInfiniteVariable] = []
        self.point_vars: List[PointVariable] = []
        self.constraints: List[ConstraintData] = []
        self.piecewise_vars: "OrderedDict[InfiniteParameter, List[InfiniteVariable]]" = OrderedDict()
        self.objective_sense = None
        self.objective_expr = None

    # -- parameters --------------------------------------------------------------------------------
    def infinite_parameter(self, lb=None, ub=None, num_supports=None, supports=None, derivative_method=None):
        if supports is None:
            supports = np.linspace(lb, ub, num_supports)
        p = InfiniteParameter(self, len(self.param_groups) + 1, 0, False, lb, ub, derivative_method)
        self.param_groups.append([p])
        self.public.append(np.unique(np.asarray(supports, dtype=np.float64)))
        self.supports.append(self.public[-1])
        return p

    def add_supports(self, pref: InfiniteParameter, values):
        g = pref.group - 1
        self.public[g] = np.unique(np.concatenate([self.public[g], np.asarray(values, dtype=np.float64)]))
        self.supports[g] = self.public[g]

    def dependent_parameters(self, supports: np.ndarray):
        """``@infinite_parameter(m, ξ[1:n] ~ dist, num_supports = K)``: one group, supports (n, K)"""
        supports = np.asarray(supports, dtype=np.float64)
        g = len(self.param_groups) + 1
        ps = [InfiniteParameter(self, g, i, True) for i in range(supports.shape[0])]
        self.param_groups.append(ps)
        self.public.append(supports)
        self.supports.append(supports)
        return ps

    def finite_parameter(self, value):
        p = FiniteParameter(self, value); self.finite_params.append(p); return p

    def parameter_function(self, func, *prefs):
        p = ParameterFunction(self, func, prefs); self.param_funcs.append(p); return p

    # -- variables ---------------------------------------------------------------------------------
    def variable(self, *prefs, lb=None, ub=None, start=None, fix=None, binary=False, integer=False):
        info = VarInfo(lb, ub, fix, start, bool(binary), bool(integer))
        if prefs:
            v = InfiniteVariable(self, prefs, info); self.infinite_vars.append(v)
        else:
            v = FiniteVariable(self, info); self.finite_vars.append(v)
        return v

    def deriv(self, var: InfiniteVariable, pref: InfiniteParameter) -> Derivative:
        for d in self.derivatives:
            if d.arg is var and d.pref is pref: return d
        d = Derivative(self, var, pref); self.derivatives.append(d); return d

    def _point(self, base, vals):
        for p in self.point_vars:
            if p.base is base and p.values == vals: return p
        p = PointVariable(self, base, vals); self.point_vars.append(p); return p

    def _semi(self, base, fixed):
        for s in self.semi_vars:
            if s.base is base and s.fixed == fixed: return s
        s = SemiInfiniteVariable(self, base, fixed); self.semi_vars.append(s); return s

    def constant_over_collocation(self, var: InfiniteVariable, pref: InfiniteParameter):
        self.piecewise_vars.setdefault(pref, []).append(var)

    # -- measures ------------------------------------------------------------------------------------
    def integral(self, expr, pref: InfiniteParameter) -> Measure:
        """``∫(expr, pref)``: trapezoid rule over ALL supports of pref (label All: internal collocation
        nodes included); the coefficient data is generated when the model is transcribed."""
        return Measure(self, expr, (pref,), None, None)

    def expect(self, expr, pref) -> Measure:
        """``𝔼(expr, ξ)``: equal weights 1/K over the supports of the (dependent) parameter group"""
        prefs = tuple(pref) if isinstance(pref, (list, tuple)) else (pref,)
        m = Measure(self, expr, prefs, None, None)
        m.kind = "expect"
        return m

    # -- constraints / objective ---------------------------------------------------------------------
    def constraint(self, expr, sense: str, rhs=0.0, restriction=None, restriction_prefs=()):
        """JuMP normalisation: everything moves left; for affine/quadratic functions the constant moves
        into the set (``y + z - t <= 42`` for ``y + z <= 42 + t``)."""
        if isinstance(sense, tuple):                    # interval: lb <= expr <= ub
            lo, hi = sense
            f = expr
        else:
            f = expr if (_is_num(rhs) and rhs == 0) else _sub(expr, rhs)
            if sense not in ("==", "<=", ">="):     # _get_constr_bounds(set) fallback, transform.jl:408-411
                raise ValueError(f"Constraint set `{sense}` is not supported by InfiniteExaModels, "
                                 "if you need support for this constraint type, please open an issue.")
            lo, hi = {"==": (0.0, 0.0), "<=": (-np.inf, 0.0), ">=": (0.0, np.inf)}[sense]
        if isinstance(f, Ref): f = _aff(f)
        if isinstance(f, AffExpr):
            c = f.constant; f = AffExpr(f.terms, 0.0); lo, hi = lo - c, hi - c
        elif isinstance(f, QuadExpr):
            c = f.aff.constant; f = QuadExpr(f.terms, AffExpr(f.aff.terms, 0.0)); lo, hi = lo - c, hi - c
        self.constraints.append(ConstraintData(f, lo, hi, restriction, tuple(restriction_prefs)))
        return self.constraints[-1]

    def objective(self, sense: str, expr):
        self.objective_sense, self.objective_expr = sense, expr

"""Expression objects of the host-side mirror of ExaModels' modelling API.

The reference builds ExaModels node objects in ``src/transform.jl:337-389`` (``_exafy``) and
``:290-334`` (``_map_variable``): ``Var`` leaves whose index is either a constant or an integer
expression over ``data_src[group_alias]`` fields, ``Parameter`` lookups, ``data_src[alias]`` float
fields, and the operators of ``src/operators.jl:2-46``.  This module provides the same vocabulary
in Python and lowers a tree to the postfix tape + affine index expressions of ``include/iexa.h``.
"""
from __future__ import annotations

import numbers
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np

# --- operator codes (must match include/iexa.h) ---------------------------------------------
OP = dict(
    CONST=0, FIELD=1, VAR=2, PAR=3,
    ADD=10, SUB=11, MUL=12, DIV=13, POW=14,
    NEG=20, POS=21, INV=22, SQRT=23, CBRT=24, ABS=25, ABS2=26, EXP=27, EXP2=28, LOG=29,
    LOG2=30, LOG10=31, LOG1P=32, SIN=33, COS=34, TAN=35, ASIN=36, ACOS=37, CSC=38, SEC=39,
    COT=40, ATAN=41, ACOT=42, SIND=43, COSD=44, TAND=45, CSCD=46, SECD=47, COTD=48, ATAND=49,
    ACOTD=50, SINH=51, COSH=52, TANH=53, CSCH=54, SECH=55, COTH=56, ATANH=57, ACOTH=58,
)
OP_NAME = {v: k for k, v in OP.items()}

# JuMP operator symbol -> tape op: the table of src/operators.jl:2-46, INCLUDING the reference's
# ``:csch => csc`` entry (operators.jl:41).  ``nl_op(sym, compat=False)`` gives the true csch.
_OP_MAPPINGS = {
    "+": "ADD", "-": "SUB", "*": "MUL", "/": "DIV", "^": "POW",
    "inv": "INV", "sqrt": "SQRT", "cbrt": "CBRT", "abs": "ABS", "abs2": "ABS2", "exp": "EXP",
    "exp2": "EXP2", "log": "LOG", "log2": "LOG2", "log10": "LOG10", "log1p": "LOG1P",
    "sin": "SIN", "cos": "COS", "tan": "TAN", "asin": "ASIN", "acos": "ACOS", "csc": "CSC",
    "sec": "SEC", "cot": "COT", "atan": "ATAN", "acot": "ACOT", "sind": "SIND", "cosd": "COSD",
    "tand": "TAND", "cscd": "CSCD", "secd": "SECD", "cotd": "COTD", "atand": "ATAND",
    "acotd": "ACOTD", "sinh": "SINH", "cosh": "COSH", "tanh": "TANH", "csch": "CSC",
    "sech": "SECH", "coth": "COTH", "atanh": "ATANH", "acoth": "ACOTH",
}


def nl_op(sym: str, compat: bool = True):
    """``_nl_op`` (src/operators.jl:49-54): operator symbol -> callable building a node."""
    if sym not in _OP_MAPPINGS:
        raise ValueError(
            f"`InfiniteExaModel`s does not support the nonlinear operator `{sym}`. "
            "If you need support for this operator, please open an issue.")
    name = _OP_MAPPINGS[sym]
    if sym == "csch" and not compat:
        name = "CSCH"
    code = OP[name]
    if 10 <= code <= 14:
        def nary(*args):
            if len(args) == 1 and code in (OP["ADD"], OP["SUB"]):
                return Unary(OP["POS"] if code == OP["ADD"] else OP["NEG"], as_node(args[0]))
            out = as_node(args[0])
            for a in args[1:]:  # Julia folds n-ary + and * left to right
                out = _binary(code, out, a)
            return out
        return nary
    return lambda a: _unary(code, a)


# --- integer index expressions ----------------------------------------------------------------
class IndexExpr:
    """Affine integer expression ``const + sum coef*int_field`` (1-based result)."""

    __slots__ = ("const", "terms")

    def __init__(self, const=0, terms=None):
        self.const = int(const)
        self.terms: Dict[str, int] = {k: int(v) for k, v in (terms or {}).items() if v != 0}

    @staticmethod
    def of(v) -> "IndexExpr":
        if isinstance(v, IndexExpr):
            return v
        if isinstance(v, DataField):
            return IndexExpr(0, {v.name: 1})
        if isinstance(v, (numbers.Integral, np.integer)):
            return IndexExpr(int(v))
        raise TypeError(f"cannot use {type(v).__name__} as an index")

    def __add__(self, o):
        o = IndexExpr.of(o)
        t = dict(self.terms)
        for k, c in o.terms.items():
            t[k] = t.get(k, 0) + c
        return IndexExpr(self.const + o.const, t)

    __radd__ = __add__

    def __neg__(self):
        return IndexExpr(-self.const, {k: -c for k, c in self.terms.items()})

    def __sub__(self, o):
        return self + (-IndexExpr.of(o))

    def __rsub__(self, o):
        return IndexExpr.of(o) + (-self)

    def __mul__(self, o):
        if not isinstance(o, (numbers.Integral, np.integer)):
            raise TypeError("index expressions are affine: multiply by integers only")
        return IndexExpr(self.const * int(o), {k: c * int(o) for k, c in self.terms.items()})

    __rmul__ = __mul__

    def key(self):
        return (self.const, tuple(sorted(self.terms.items())))

    def __repr__(self):
        return "Idx(" + " + ".join([str(self.const)] + [f"{c}*{k}" for k, c in self.terms.items()]) + ")"


# --- value nodes ---------------------------------------------------------------------------------
class Node:
    __array_priority__ = 1000

    def __add__(self, o): return _binary(OP["ADD"], self, o)
    def __radd__(self, o): return _binary(OP["ADD"], o, self)
    def __sub__(self, o): return _binary(OP["SUB"], self, o)
    def __rsub__(self, o): return _binary(OP["SUB"], o, self)
    def __mul__(self, o): return _binary(OP["MUL"], self, o)
    def __rmul__(self, o): return _binary(OP["MUL"], o, self)
    def __truediv__(self, o): return _binary(OP["DIV"], self, o)
    def __rtruediv__(self, o): return _binary(OP["DIV"], o, self)
    def __pow__(self, o): return _binary(OP["POW"], self, o)
    def __rpow__(self, o): return _binary(OP["POW"], o, self)
    def __neg__(self): return Unary(OP["NEG"], self)
    def __pos__(self): return Unary(OP["POS"], self)


@dataclass(eq=False)
class Const(Node):
    """Literal; a whole expression that is a literal is ``ExaModels.Null(c)`` (transform.jl:393)."""
    c: float


Null = Const


@dataclass(eq=False)
class DataField(Node):
    """``data_src[alias]``: float field in value position (transform.jl:320-322) or integer field
    in index position (transform.jl:309,317)."""
    name: str

    # index arithmetic such as ``idx - 1`` (make_reduced_expr, transform.jl:485-505)
    def idx(self) -> IndexExpr:
        return IndexExpr.of(self)


@dataclass(eq=False)
class Var(Node):
    index: IndexExpr


@dataclass(eq=False)
class Par(Node):
    index: IndexExpr


@dataclass(eq=False)
class Unary(Node):
    op: int
    a: Node


@dataclass(eq=False)
class Binary(Node):
    op: int
    a: Node
    b: Node


def as_node(v) -> Node:
    if isinstance(v, Node):
        return v
    if isinstance(v, (numbers.Real, np.floating, np.integer)):
        return Const(float(v))
    raise TypeError(f"cannot convert {type(v).__name__} to an expression node")


def _is_num(v):
    return isinstance(v, (numbers.Real, np.floating, np.integer)) and not isinstance(v, Node)


def _binary(op, a, b):
    if _is_num(a) and _is_num(b):
        raise TypeError("both operands are plain numbers")
    return Binary(op, as_node(a), as_node(b))


def _unary(op, a):
    return Unary(op, as_node(a))


class DataSource:
    """``ExaModels.DataSource()`` (transform.jl:453): ``ds[alias]`` / ``ds.alias``."""

    def __getitem__(self, name: str) -> DataField:
        return DataField(str(name))

    def __getattr__(self, name: str) -> DataField:
        if name.startswith("__"):
            raise AttributeError(name)
        return DataField(name)


# math functions named like Julia's
def _mk(name):
    code = OP[name]
    return lambda a: _unary(code, a)


inv, sqrt, cbrt, abs_, abs2, exp, exp2, log, log2, log10, log1p = map(
    _mk, ["INV", "SQRT", "CBRT", "ABS", "ABS2", "EXP", "EXP2", "LOG", "LOG2", "LOG10", "LOG1P"])
sin, cos, tan, asin, acos, csc, sec, cot, atan, acot = map(
    _mk, ["SIN", "COS", "TAN", "ASIN", "ACOS", "CSC", "SEC", "COT", "ATAN", "ACOT"])
sind, cosd, tand, cscd, secd, cotd, atand, acotd = map(
    _mk, ["SIND", "COSD", "TAND", "CSCD", "SECD", "COTD", "ATAND", "ACOTD"])
sinh, cosh, tanh, csch, sech, coth, atanh, acoth = map(
    _mk, ["SINH", "COSH", "TANH", "CSCH", "SECH", "COTH", "ATANH", "ACOTH"])


# --- lowering to the C-ABI tape --------------------------------------------------------------
NODE_DTYPE = np.dtype([("op", "<i4"), ("a", "<i4"), ("b", "<i4"), ("pad", "<i4"), ("c", "<f8")])
INDEX_DTYPE = np.dtype([("base", "<i8"), ("nterms", "<i4"), ("col", "<i4", (4,)), ("pad", "<i4"),
                        ("coef", "<i8", (4,))])
assert NODE_DTYPE.itemsize == 24 and INDEX_DTYPE.itemsize == 64


@dataclass
class Tape:
    nodes: np.ndarray   # NODE_DTYPE
    index: np.ndarray   # INDEX_DTYPE
    n_var_leaves: int = 0


def lower(expr, int_cols: List[str], fp_cols: List[str]) -> Tape:
    """Postfix tape of ``expr`` against an iterator with the given column names.

    Iterative post-order (expanded-measure trees are O(K) deep, transform.jl:430-435)."""
    expr = as_node(expr)
    icol = {n: i for i, n in enumerate(int_cols)}
    fcol = {n: i for i, n in enumerate(fp_cols)}
    nodes: List[Tuple[int, int, int, float]] = []
    idx_rows: List[Tuple[int, List[Tuple[int, int]]]] = []
    idx_ids: Dict[int, int] = {}
    nvar = 0

    def index_id(ix: IndexExpr) -> int:
        # every Var/Par leaf gets its own index record; the engine canonicalises/deduplicates
        terms = []
        for name, coef in ix.terms.items():
            if name not in icol:
                raise KeyError(f"index field `{name}` is not an integer column of the iterator "
                               f"(have {int_cols})")
            terms.append((icol[name], coef))
        if len(terms) > 4:
            raise ValueError("index expression with more than 4 integer fields")
        idx_rows.append((ix.const, terms))
        return len(idx_rows) - 1

    out_id: Dict[int, int] = {}
    stack: List[Tuple[Node, bool]] = [(expr, False)]
    while stack:
        nd, done = stack.pop()
        if id(nd) in out_id:  # shared sub-expression object: emit once, reference twice
            continue
        if not done:
            if isinstance(nd, Binary):
                stack.append((nd, True)); stack.append((nd.b, False)); stack.append((nd.a, False))
                continue
            if isinstance(nd, Unary):
                stack.append((nd, True)); stack.append((nd.a, False))
                continue
        if isinstance(nd, Const):
            nodes.append((OP["CONST"], 0, 0, float(nd.c)))
        elif isinstance(nd, DataField):
            if nd.name not in fcol:
                raise KeyError(f"value field `{nd.name}` is not a float column of the iterator "
                               f"(have {fp_cols})")
            nodes.append((OP["FIELD"], fcol[nd.name], 0, 0.0))
        elif isinstance(nd, Var):
            nodes.append((OP["VAR"], index_id(nd.index), 0, 0.0)); nvar += 1
        elif isinstance(nd, Par):
            nodes.append((OP["PAR"], index_id(nd.index), 0, 0.0))
        elif isinstance(nd, Unary):
            nodes.append((nd.op, out_id[id(nd.a)], 0, 0.0))
        elif isinstance(nd, Binary):
            nodes.append((nd.op, out_id[id(nd.a)], out_id[id(nd.b)], 0.0))
        else:
            raise TypeError(f"unknown node {nd!r}")
        out_id[id(nd)] = len(nodes) - 1
    arr = np.zeros(len(nodes), dtype=NODE_DTYPE)
    for i, (op, a, b, c) in enumerate(nodes):
        arr[i] = (op, a, b, 0, c)
    ix = np.zeros(max(len(idx_rows), 1), dtype=INDEX_DTYPE)
    for i, (base, terms) in enumerate(idx_rows):
        ix[i]["base"] = base
        ix[i]["nterms"] = len(terms)
        for t, (c, coef) in enumerate(terms):
            ix[i]["col"][t] = c
            ix[i]["coef"][t] = coef
    return Tape(arr, ix[: len(idx_rows)] if idx_rows else ix[:0], nvar)


def eval_numpy(tape: Tape, fp_cols, theta, K: int) -> np.ndarray:
    """value of a tape without Var leaves at every support, with numpy (the HOST statement of a parameter function —
    transform.jl:161-183 calls the Julia closure per support; the engine evaluates the same tape on the device).
    PAR leaves with constant indices read ``theta``."""
    D2R = np.pi / 180.0
    un = {OP["NEG"]: np.negative, OP["POS"]: lambda u: u, OP["INV"]: lambda u: 1.0 / u, OP["SQRT"]: np.sqrt, OP["CBRT"]: np.cbrt,
          OP["ABS"]: np.abs, OP["ABS2"]: lambda u: u * u, OP["EXP"]: np.exp, OP["EXP2"]: np.exp2, OP["LOG"]: np.log,
          OP["LOG2"]: np.log2, OP["LOG10"]: np.log10, OP["LOG1P"]: np.log1p, OP["SIN"]: np.sin, OP["COS"]: np.cos,
          OP["TAN"]: np.tan, OP["ASIN"]: np.arcsin, OP["ACOS"]: np.arccos, OP["ATAN"]: np.arctan,
          OP["SINH"]: np.sinh, OP["COSH"]: np.cosh, OP["TANH"]: np.tanh, OP["ATANH"]: np.arctanh,
          OP["SIND"]: lambda u: np.sin(u * D2R), OP["COSD"]: lambda u: np.cos(u * D2R), OP["TAND"]: lambda u: np.tan(u * D2R)}
    vals = []
    for nd in tape.nodes:
        op, a, b, c = int(nd["op"]), int(nd["a"]), int(nd["b"]), float(nd["c"])
        if op == OP["CONST"]:
            v = np.full(K, c)
        elif op == OP["FIELD"]:
            v = np.asarray(fp_cols[a], dtype=np.float64)
        elif op == OP["PAR"]:
            ix = tape.index[a]
            if int(ix["nterms"]) != 0:
                raise ValueError("parameter functions may reference finite parameters only")
            v = np.full(K, float(theta[int(ix["base"]) - 1]))
        elif op == OP["VAR"]:
            raise ValueError("parameter functions cannot reference variables")
        elif op == OP["ADD"]:
            v = vals[a] + vals[b]
        elif op == OP["SUB"]:
            v = vals[a] - vals[b]
        elif op == OP["MUL"]:
            v = vals[a] * vals[b]
        elif op == OP["DIV"]:
            v = vals[a] / vals[b]
        elif op == OP["POW"]:
            v = vals[a] * vals[a] if np.all(vals[b] == 2.0) else np.power(vals[a], vals[b])
        elif op in un:
            v = un[op](vals[a])
        else:
            raise ValueError(f"eval_numpy: operator {op} not supported in parameter functions")
        vals.append(v)
    return np.ascontiguousarray(vals[-1], dtype=np.float64)

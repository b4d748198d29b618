"""Python restatement of ``src/transform.jl``: InfiniteModel -> ExaCore lowering.

Function names, argument meaning, emission ORDER and error behaviour follow the reference (cited
per function).  The input is the small modelling layer of ``infopt.py``; the output is the
``ExaCore`` of ``core.py``, i.e. exactly what the reference registers with ExaModels and what the
engine's C ABI consumes.  The InfiniteOpt pieces the reference *calls into* (supports, measure
coefficients, derivative templates) are restated here from InfiniteOpt's documented behaviour:
trapezoid quadrature over all supports, expectation weights 1/K, backward finite difference
``Δt·d[i] − y[i] + y[i−1] = 0`` and Lobatto orthogonal collocation.
"""
from __future__ import annotations

import itertools
import warnings
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import expr as E
from . import infopt as io
from .core import Constraint, ExaCore, Itr, Parameter, Variable
from .expr import DataSource, IndexExpr, nl_op

_ObjMeasureExpansionWarn = (
    "Unable to convert objective measures into a form that is efficient for ExaModels using existing "
    "heuristics. Performance may be significantly degraded. Try simplying the objective structure. "
    "if you think this form should be supported, please open an issue.")


@dataclass
class ExaMappingData:
    """``ExaMappingData`` (src/infiniteopt_backend.jl:12-57)."""
    infvar_mappings: Dict = field(default_factory=dict)
    finvar_mappings: Dict = field(default_factory=dict)      # ref -> 1-based x index
    param_mappings: Dict = field(default_factory=dict)       # ref -> Parameter block
    constraint_mappings: Dict = field(default_factory=dict)
    param_alias: Dict = field(default_factory=dict)          # pref -> float field name
    group_alias: List[str] = field(default_factory=list)     # group -> integer field name
    base_itrs: List[Itr] = field(default_factory=list)
    support_to_index: Dict = field(default_factory=dict)     # (group, support) -> index
    support_labels: List = field(default_factory=list)
    deriv_data_cache: Dict = field(default_factory=dict)
    has_internal_supps: List[bool] = field(default_factory=list)
    semivar_info: Dict = field(default_factory=dict)
    supports: List[np.ndarray] = field(default_factory=list)  # per group, after generative supports
    internal: List[np.ndarray] = field(default_factory=list)  # per group: mask of internal (collocation) supports


# ---- transform.jl:2-38 ----------------------------------------------------------------------------
def _build_base_iterators(data: ExaMappingData, inf_model: io.InfiniteModel) -> None:
    for g, group in enumerate(inf_model.param_groups, start=1):
        dependent = group[0].dependent
        for pref in group:
            data.param_alias[pref] = f"dp{g}{pref.pos + 1}" if dependent else f"ip{g}"
        aliases = [data.param_alias[p] for p in group]
        itr_sym = f"group_idx{len(data.group_alias) + 1}"
        data.group_alias.append(itr_sym)
        # add_generative_supports (transform.jl:22): internal collocation nodes
        supps = np.asarray(inf_model.public[g - 1], dtype=np.float64)
        internal = np.zeros(supps.shape[-1], dtype=bool)
        method = group[0].derivative_method
        if not dependent and isinstance(method, io.OrthogonalCollocation) and method.num_nodes > 2:
            supps, internal = _add_generative_supports(supps, method.num_nodes)
        inf_model.supports[g - 1] = supps
        data.supports.append(supps)
        data.internal.append(internal)
        K = supps.shape[-1]
        # support_to_index[(g, s)] = i (transform.jl:27-29) is answered on demand by _nearest_index (a binary search in
        # the sorted supports of an independent parameter) instead of a dictionary with one entry per support
        if dependent:
            fps = OrderedDict((a, supps[j]) for j, a in enumerate(aliases))
        else:
            fps = OrderedDict([(aliases[0], supps)])
        data.base_itrs.append(Itr(K, {itr_sym: np.arange(1, K + 1)}, fps))
        data.support_labels.append(internal)          # True = internal collocation node, False = public support
        data.has_internal_supps.append(bool(internal.any()))


def _lobatto_nodes(n: int) -> np.ndarray:
    """Gauss-Lobatto nodes on [-1, 1] (end points + roots of P'_{n-1})"""
    if n == 2:
        return np.array([-1.0, 1.0])
    inner = np.polynomial.legendre.Legendre.basis(n - 1).deriv().roots()
    return np.concatenate([[-1.0], np.sort(inner.real), [1.0]])


def _add_generative_supports(pub: np.ndarray, num_nodes: int):
    nodes = _lobatto_nodes(num_nodes)[1:-1]
    n = len(nodes)
    lo, hi = pub[:-1], pub[1:]
    T = len(pub) + n * (len(pub) - 1)
    out = np.empty(T)
    internal = np.ones(T, dtype=bool)
    out[0::n + 1] = pub
    internal[0::n + 1] = False
    for j, z in enumerate(nodes):
        out[1 + j::n + 1] = lo + (z + 1) / 2 * (hi - lo)
    return out, internal


# ---- transform.jl:41-101 ---------------------------------------------------------------------------
def _process_value(val, supp):
    return val(*supp) if callable(val) else val


def _ensure_continuous(info: io.VarInfo):
    """transform.jl:41-45"""
    if info.binary or info.integer:
        raise ValueError("Integer variables are not supported by ExaModels.")


def _get_variable_bounds_and_start(info: io.VarInfo, itrs: Optional[List[Itr]] = None, data=None, groups=()):
    _ensure_continuous(info)
    vals = (info.lb, info.ub, info.fix, info.start)
    if itrs is None or not any(callable(v) for v in vals):
        lb, ub, start = -np.inf, np.inf, 0.0
        if info.fix is not None: lb = ub = info.fix
        if info.lb is not None: lb = info.lb
        if info.ub is not None: ub = info.ub
        if info.start is not None: start = info.start
        return lb, ub, start
    dims = tuple(it.K for it in itrs)
    lb, ub, start = np.full(dims, -np.inf), np.full(dims, np.inf), np.zeros(dims)
    for idx in itertools.product(*[range(d) for d in dims]):
        supp = []
        for g, i in zip(groups, idx):
            s = data.supports[g - 1]
            supp.extend([s[i]] if s.ndim == 1 else list(s[:, i]))
        if info.fix is not None: lb[idx] = ub[idx] = _process_value(info.fix, supp)
        if info.lb is not None: lb[idx] = _process_value(info.lb, supp)
        if info.ub is not None: ub[idx] = _process_value(info.ub, supp)
        if info.start is not None: start[idx] = _process_value(info.start, supp)
    return lb, ub, start


# ---- transform.jl:104-183 ----------------------------------------------------------------------------
def _add_finite_variables(core: ExaCore, data, inf_model):
    for vref in inf_model.finite_vars:
        lb, ub, start = _get_variable_bounds_and_start(vref.info)
        new_var = core.add_var(1, start=start, lvar=lb, uvar=ub)
        data.finvar_mappings[vref] = new_var.index(1)
    return core


def _add_finite_parameters(core: ExaCore, data, inf_model):
    for pref in inf_model.finite_params:
        data.param_mappings[pref] = core.add_par([pref.value])
    return core


def _add_infinite_variables(core: ExaCore, data, inf_model):
    for vref in list(inf_model.infinite_vars) + list(inf_model.derivatives):
        group_idxs = vref.groups
        itrs = [data.base_itrs[g - 1] for g in group_idxs]
        lb, ub, start = _get_variable_bounds_and_start(vref.info, itrs, data, group_idxs)
        dims = tuple(it.K for it in itrs)
        data.infvar_mappings[vref] = core.add_var(*dims, start=start, lvar=lb, uvar=ub)
    return core


def _support_values(data, groups, idx):
    supp = []
    for g, i in zip(groups, idx):
        s = data.supports[g - 1]
        supp.extend([s[i]] if s.ndim == 1 else list(s[:, i]))
    return supp


def _evaluate_parameter_function(func, data, group_idxs, dims) -> np.ndarray:
    """table of a parameter function at every support combination (transform.jl:174-177 calls the closure once per
    combination).  First try ONE call on broadcast support arrays — exact for functions written with numpy ufuncs —
    and verify it against per-combination calls on a few entries; fall back to the per-combination loop otherwise
    (e.g. functions that branch on their arguments, test/solve.jl:99-105)."""
    if all(data.supports[g - 1].ndim == 1 for g in group_idxs) and int(np.prod(dims)) > 64:
        try:
            grids = np.meshgrid(*[data.supports[g - 1] for g in group_idxs], indexing="ij")
            vals = np.asarray(func(*grids), dtype=np.float64)
            if vals.shape == tuple(dims):
                rng = np.random.default_rng(0)
                probes = [tuple(int(rng.integers(0, d)) for d in dims) for _ in range(8)] + [tuple(0 for _ in dims), tuple(d - 1 for d in dims)]
                if all(vals[i] == float(func(*_support_values(data, group_idxs, i))) for i in probes):
                    return vals
        except Exception:
            pass
    vals = np.empty(dims)
    for idx in itertools.product(*[range(d) for d in dims]):
        vals[idx] = func(*_support_values(data, group_idxs, idx))
    return vals


def _add_parameter_functions(core: ExaCore, data, inf_model):
    for pfref in inf_model.param_funcs:
        group_idxs = pfref.groups
        dims = tuple(data.base_itrs[g - 1].K for g in group_idxs)
        data.param_mappings[pfref] = core.add_par(_evaluate_parameter_function(pfref.func, data, group_idxs, dims))
    return core


# ---- transform.jl:186-287 --------------------------------------------------------------------------------
def _nearest_index(data, g, value):
    """``data.support_to_index[g, value]`` (transform.jl:27-29, used by :199,:202,:266)"""
    s = data.supports[g - 1]
    if s.ndim == 1 and len(s) > 1:                   # supports of an independent parameter come sorted (transform.jl:528)
        j = int(np.searchsorted(s, value))
        cand = [c for c in (j - 1, j) if 0 <= c < len(s)]
        i = min(cand, key=lambda c: abs(s[c] - value))
    else:
        i = int(np.argmin(np.abs(s - value))) if s.ndim == 1 else int(np.argmin(np.abs(s - np.asarray(value).reshape(-1, 1)).sum(axis=0)))
    ref = s[i] if s.ndim == 1 else s[:, i]
    if np.max(np.abs(ref - value)) > 1e-12 * max(1.0, float(np.max(np.abs(value)))):
        raise KeyError(f"support {value} of parameter group {g} does not exist")
    return i + 1


def _process_semi_infinite_var(vref: io.SemiInfiniteVariable, data):
    ivref = vref.base
    indexing = []
    for i, p in enumerate(ivref.prefs):
        if i in vref.fixed:
            indexing.append(_nearest_index(data, p.group, vref.fixed[i]))     # store the support index
        else:
            indexing.append(data.group_alias[p.group - 1])                      # group alias for indexing
    mapped = data.param_mappings[ivref] if isinstance(ivref, io.ParameterFunction) else data.infvar_mappings[ivref]
    data.semivar_info[vref] = (mapped, indexing)
    return data.semivar_info[vref]


def _update_bounds_and_start(core: ExaCore, info: io.VarInfo, i: int):
    if info.lb is not None: core.lvar_vec[i - 1] = info.lb
    if info.ub is not None: core.uvar_vec[i - 1] = info.ub
    if info.fix is not None: core.lvar_vec[i - 1] = core.uvar_vec[i - 1] = info.fix
    if info.start is not None: core.x0_vec[i - 1] = info.start


def _add_semi_infinite_variables(core, data, inf_model):
    for vref in inf_model.semi_vars:
        mapped, indexing = _process_semi_infinite_var(vref, data)
        info = vref.info
        if any(v is not None for v in (info.lb, info.ub, info.fix, info.start)):
            ranges = [[ix] if isinstance(ix, int) else range(1, mapped.size[i] + 1) for i, ix in enumerate(indexing)]
            for idx in itertools.product(*ranges):
                _update_bounds_and_start(core, info, mapped.index(*idx))


def _process_point_var(vref: io.PointVariable, data) -> int:
    ivref = vref.base
    idxs = [_nearest_index(data, p.group, v) for p, v in zip(ivref.prefs, vref.values)]
    return data.infvar_mappings[ivref].index(*idxs)


def _add_point_variables(core, data, inf_model):
    for vref in inf_model.point_vars:
        pt = _process_point_var(vref, data)
        data.finvar_mappings[vref] = pt
        _update_bounds_and_start(core, vref.info, pt)


# ---- transform.jl:290-334 ------------------------------------------------------------------------------------
def _map_variable(vref, data_src: DataSource, data: ExaMappingData):
    t = vref.index_type
    if t == "FiniteVariable":
        return E.Var(IndexExpr(data.finvar_mappings[vref]))
    if t == "PointVariable":
        if vref not in data.finvar_mappings:
            data.finvar_mappings[vref] = _process_point_var(vref, data)
        return E.Var(IndexExpr(data.finvar_mappings[vref]))
    if t in ("InfiniteVariable", "Derivative"):
        idx_pars = tuple(data_src[data.group_alias[g - 1]] for g in vref.groups)
        return data.infvar_mappings[vref][idx_pars]
    if t == "SemiInfiniteVariable":
        if vref not in data.semivar_info:
            _process_semi_infinite_var(vref, data)
        ivar, inds = data.semivar_info[vref]
        return ivar[tuple(i if isinstance(i, int) else data_src[i] for i in inds)]
    if t in ("IndependentParameter", "DependentParameter"):
        return data_src[data.param_alias[vref]]
    if t == "FiniteParameter":
        return data.param_mappings[vref][1]
    if t == "ParameterFunction":
        idx_pars = tuple(data_src[data.group_alias[g - 1]] for g in vref.groups)
        return data.param_mappings[vref][idx_pars]
    raise ValueError(f"Unable to add `{vref}` to an ExaModel, it's index type `{t}` is not yet supported by "
                     "InfiniteExaModels.")


# ---- transform.jl:337-393 -------------------------------------------------------------------------------------
def _isone(c): return c == 1.0


def _exafy(ex, data_src, data):
    if isinstance(ex, io.Ref):
        return _map_variable(ex, data_src, data)
    if io._is_num(ex):
        return float(ex)
    if isinstance(ex, io.AffExpr):
        c = ex.constant
        if ex.terms:
            out = None
            for v, coef in ex.terms.items():
                v_ex = _map_variable(v, data_src, data)
                term = v_ex if _isone(coef) else coef * v_ex
                out = term if out is None else out + term
            return out if c == 0 else out + c
        return c
    if isinstance(ex, io.QuadExpr):
        aff = _exafy(ex.aff, data_src, data)
        if ex.terms:
            out = None
            for (v1, v2), coef in ex.terms.items():
                if v1 is v2:
                    v_ex = _map_variable(v1, data_src, data)
                    term = E.abs2(v_ex) if _isone(coef) else coef * E.abs2(v_ex)
                else:
                    a, b = _map_variable(v1, data_src, data), _map_variable(v2, data_src, data)
                    term = a * b if _isone(coef) else coef * a * b
                out = term if out is None else out + term
            is_zero_aff = not ex.aff.terms and ex.aff.constant == 0
            return out if is_zero_aff else out + aff
        return aff
    if isinstance(ex, io.NLExpr):
        return nl_op(ex.head)(*[_exafy(a, data_src, data) for a in ex.args])
    raise TypeError(f"cannot convert {type(ex).__name__}")


def _finalize_expr(ex):
    return E.Null(float(ex)) if io._is_num(ex) else ex


# ---- transform.jl:396-462 --------------------------------------------------------------------------------------
def _support_in_restriction(restriction, prefs, row: Dict[str, float], data) -> bool:
    return bool(restriction(*[row[data.param_alias[p]] for p in prefs]))


def _product(itrs: List[Itr]) -> Itr:
    return itrs[0] if len(itrs) == 1 else Itr.product(itrs)


def _add_constraints(core: ExaCore, data, inf_model):
    for constr in inf_model.constraints:
        expr = constr.expr
        if any(v.index_type == "Measure" for v in io.all_expression_variables(expr)):
            warnings.warn("Constrained measures can lead to poor performance with ExaModels.")
            expr = expand_measures(expr, data)
        group_idxs = io.expression_groups(expr)
        if not group_idxs:
            itr = Itr.empty()
        else:
            itr = _product([data.base_itrs[g - 1] for g in group_idxs])
        if constr.restriction is not None:
            ic, fc = itr.materialise()
            names = itr.fp_names()
            mask = np.array([_support_in_restriction(constr.restriction, constr.restriction_prefs,
                                                     {n: c[k] for n, c in zip(names, fc)}, data) for k in range(itr.K)], dtype=bool)
            itr = itr.filtered(mask)
        data_src = DataSource()
        em_expr = _finalize_expr(_exafy(expr, data_src, data))
        con = core.add_con(em_expr, itr, lcon=constr.lb, ucon=constr.ub)
        data.constraint_mappings[id(constr)] = con
    return core


# ---- transform.jl:465-562 ---------------------------------------------------------------------------------------
def make_reduced_expr(vref, pref, idx, data_src, data):
    """index of ``vref`` with the operator parameter's group replaced by the index expression ``idx``"""
    alias = data.group_alias[pref.group - 1]
    if vref.index_type == "SemiInfiniteVariable":
        ivar, inds = data.semivar_info[vref]
        return ivar[tuple(i if isinstance(i, int) else (idx if i == alias else data_src[i]) for i in inds)]
    idx_pars = tuple(idx if data.group_alias[g - 1] == alias else data_src[data.group_alias[g - 1]] for g in vref.groups)
    return data.infvar_mappings[vref][idx_pars]


def derivative_expr_data(pref, supps: np.ndarray, internal: np.ndarray, method):
    """row iterator columns of the derivative approximation along ``pref`` (InfiniteOpt's
    ``derivative_expr_data``): returns (int columns, fp columns)"""
    T = len(supps)
    if isinstance(method, io.FiniteDifference):
        if method.kind == "backward":
            idxs = np.arange(2, T + 1)
            return {"idx": idxs}, {"d_arg1": supps[idxs - 1] - supps[idxs - 2]}
        if method.kind == "forward":
            idxs = np.arange(1, T)
            return {"idx": idxs}, {"d_arg1": supps[idxs] - supps[idxs - 1]}
        if method.kind == "central":
            idxs = np.arange(2, T)
            return {"idx": idxs}, {"d_arg1": supps[idxs] - supps[idxs - 2]}
        raise ValueError(method.kind)
    # orthogonal collocation: per interval [lb, ub] with nodes t_1..t_n (internal nodes + ub):
    #   y(t_j) − y(lb) = Σ_k M[j,k]·dy(t_k),  M = M2·inv(M1),  M1[j,k] = k (t_j−lb)^(k−1),  M2[j,k] = (t_j−lb)^k
    n = method.num_nodes - 1
    pub = np.flatnonzero(~internal)
    a = pub[:-1]                                                     # 0-based position of every interval's lower bound
    if len(a) == 0:
        return {"idx": np.zeros(0, dtype=np.int64), "d_lb": np.zeros(0, dtype=np.int64)}, {f"d_arg{kk + 1}": np.zeros(0) for kk in range(n)}
    assert np.all(np.diff(pub) == n), "every interval carries the same number of collocation nodes"
    # all intervals at once (the reference loops over supports; at 10^6 supports that loop is the build time)
    tj = supps[a[:, None] + 1 + np.arange(n)[None, :]] - supps[a][:, None]          # (intervals, n)
    k = np.arange(1, n + 1)
    M1 = k[None, None, :] * tj[:, :, None] ** (k[None, None, :] - 1)
    M2 = tj[:, :, None] ** k[None, None, :]
    M = M2 @ np.linalg.inv(M1)                                                      # (intervals, n, n)
    lbs = np.repeat(a + 1, n)
    nodes = (a[:, None] + 2 + np.arange(n)[None, :]).reshape(-1)
    return {"idx": nodes, "d_lb": lbs}, {f"d_arg{kk + 1}": np.ascontiguousarray(M[:, :, kk].reshape(-1)) for kk in range(n)}


def make_indexed_derivative_expr(dref, vref, pref, data_src, data, method, group_alias):
    idx = data_src[group_alias].idx()
    if isinstance(method, io.FiniteDifference):
        d = lambda i: make_reduced_expr(dref, pref, i, data_src, data)
        y = lambda i: make_reduced_expr(vref, pref, i, data_src, data)
        if method.kind == "backward": return data_src.d_arg1 * d(idx) - y(idx) + y(idx - 1)
        if method.kind == "forward": return data_src.d_arg1 * d(idx) - y(idx + 1) + y(idx)
        return data_src.d_arg1 * d(idx) - y(idx + 1) + y(idx - 1)
    lb = data_src.d_lb.idx()
    n = method.num_nodes - 1
    out = None
    for k in range(n):
        term = data_src[f"d_arg{k + 1}"] * make_reduced_expr(dref, pref, lb + (k + 1), data_src, data)
        out = term if out is None else out + term
    return out - make_reduced_expr(vref, pref, idx, data_src, data) + make_reduced_expr(vref, pref, lb, data_src, data)


def _add_derivative_approximations(core: ExaCore, data, inf_model):
    for dref in inf_model.derivatives:
        vref, pref = dref.arg, dref.pref
        method = pref.derivative_method
        group_idxs = vref.groups
        pref_group = pref.group
        base_itr = data.base_itrs[pref_group - 1]
        supps = data.supports[pref_group - 1]
        if (pref, id(method)) not in data.deriv_data_cache:   # the same rows for every variable differentiated along pref
            data.deriv_data_cache[(pref, id(method))] = derivative_expr_data(pref, supps, data.internal[pref_group - 1], method)
        ints, fps = (dict(d) for d in data.deriv_data_cache[(pref, id(method))])
        galias = data.group_alias[pref_group - 1]
        idxs = ints.pop("idx")
        cols_i = OrderedDict([(galias, idxs)]); cols_i.update(ints)
        cols_f = OrderedDict((n, c[idxs - 1]) for n, c in base_itr.fps.items()); cols_f.update(fps)
        pref_itr = Itr(len(idxs), cols_i, cols_f)
        itr = _product([pref_itr if g == pref_group else data.base_itrs[g - 1] for g in group_idxs])
        data_src = DataSource()
        em_expr = make_indexed_derivative_expr(dref, vref, pref, data_src, data, method, galias)
        core.add_con(em_expr, itr)
    return core


# ---- transform.jl:565-601 ------------------------------------------------------------------------------------------
def _add_collocation_restrictions(core: ExaCore, data, inf_model):
    for pref, vrefs in inf_model.piecewise_vars.items():
        g = pref.group
        if not data.has_internal_supps[g - 1]:
            continue
        num_nodes = pref.derivative_method.num_nodes - 2
        num_supps = len(data.supports[g - 1])
        ubs = np.repeat(np.arange(2 + num_nodes, num_supps + 1, num_nodes + 1), num_nodes)
        pts = np.setdiff1d(np.arange(2, num_supps), ubs)
        pref_itr = Itr(len(ubs), {"i1": ubs, "i2": pts}, {})
        pref_alias = data.group_alias[g - 1]
        for vref in vrefs:
            aliases = [data.group_alias[gg - 1] for gg in vref.groups]
            itr = _product([pref_itr if gg == g else data.base_itrs[gg - 1] for gg in vref.groups])
            data_src = DataSource()
            ivar = data.infvar_mappings[vref]
            e1 = ivar[tuple(data_src.i1 if a == pref_alias else data_src[a] for a in aliases)]
            e2 = ivar[tuple(data_src.i2 if a == pref_alias else data_src[a] for a in aliases)]
            core.add_con(e1 - e2, itr)
    return core


# ---- transform.jl:604-767 ----------------------------------------------------------------------------------------------
def _measure_data(mref: io.Measure, data):
    g = mref.group
    supps = data.supports[g - 1]
    if getattr(mref, "kind", "integral") == "expect":
        K = supps.shape[-1]
        return supps, np.full(K, 1.0 / K)
    from .models import trapezoid_coeffs
    return supps, trapezoid_coeffs(supps)


def _make_measure_itr(mref: io.Measure, data) -> Dict[str, np.ndarray]:
    """rows of the measure iterator as named columns: c, group alias (support index), parameter aliases"""
    supps, coeffs = _measure_data(mref, data)
    g = mref.group
    alias = data.group_alias[g - 1]
    assert len(mref.prefs) == len(inf_group := mref.model.param_groups[g - 1]), \
        "we don't allow partially measured dependent parameters"
    K = supps.shape[-1]
    cols = OrderedDict([("c", coeffs), (alias, np.arange(1, K + 1))])
    for p in inf_group:
        cols[data.param_alias[p]] = supps if supps.ndim == 1 else supps[p.pos]
    return cols


def _terms_can_be_moved_inside_measure(ex, mref) -> bool:
    if isinstance(ex, (io.Ref, io.AffExpr)):
        return True
    if isinstance(ex, io.QuadExpr):
        return (mref, mref) not in ex.terms
    if isinstance(ex, io.NLExpr):
        m_inds = [a for a in ex.args if not io._is_num(a) and mref in io.all_expression_variables(a)]
        if ex.head in ("+", "-"):
            return all(_terms_can_be_moved_inside_measure(a, mref) for a in m_inds)
        if ex.head == "*":
            return len(m_inds) <= 1 and all(_terms_can_be_moved_inside_measure(a, mref) for a in m_inds)
        return False
    return False


def _product_cols(curr: Dict[str, np.ndarray], prev: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """``[(i[1]..., i[2]..., c = i[1].c * i[2].c) for i in Iterators.product(curr_itr, prev_itr)]`` (curr fastest)"""
    kc, kp = len(curr["c"]), len(prev["c"])
    out = OrderedDict()
    for n, c in curr.items(): out[n] = np.tile(c, kp)
    for n, c in prev.items(): out[n] = np.repeat(c, kc)      # later names win, like NamedTuple merging
    out["c"] = np.tile(curr["c"], kp) * np.repeat(prev["c"], kc)
    return out


def _process_measure_sum(vref: io.Measure, data, prev=None):
    mexpr = vref.expr
    curr = _make_measure_itr(vref, data)
    itr = curr if prev is None else _product_cols(curr, prev)
    mrefs = [v for v in io.all_expression_variables(mexpr) if v.index_type == "Measure"]
    if not mrefs:
        return mexpr, itr
    if len(mrefs) == 1 and _terms_can_be_moved_inside_measure(mexpr, mrefs[0]):
        mref = mrefs[0]
        inner, new_itr = _process_measure_sum(mref, data, itr)
        return io.map_expression(lambda v: inner if v is mref else v, mexpr), new_itr
    warnings.warn(_ObjMeasureExpansionWarn)
    return expand_measures(mexpr, data), itr


def _cols_to_itr(cols: Dict[str, np.ndarray]) -> Itr:
    ints = OrderedDict((n, c) for n, c in cols.items() if np.issubdtype(np.asarray(c).dtype, np.integer))
    fps = OrderedDict((n, c) for n, c in cols.items() if n not in ints)
    return Itr(len(cols["c"]), ints, fps)


def _add_generic_objective_term(core, ex, data):
    em_expr = _finalize_expr(_exafy(ex, DataSource(), data))
    core.add_obj(em_expr, Itr.empty())
    return core


def _add_objective_aff_term(core, coef, vref, data):
    if vref.index_type == "Measure":
        mexpr, cols = _process_measure_sum(vref, data)
        data_src = DataSource()
        em_expr = data_src.c * _exafy(coef * mexpr, data_src, data)
        core.add_obj(_finalize_expr(em_expr), _cols_to_itr(cols))
        return core
    return _add_generic_objective_term(core, coef * vref, data)


def _add_objective(core, ex, data, inf_model):
    if isinstance(ex, io.Ref):
        return _add_objective_aff_term(core, 1.0, ex, data)
    if isinstance(ex, io.AffExpr):
        for vref, coef in ex.terms.items():
            core = _add_objective_aff_term(core, coef, vref, data)
        if ex.constant != 0:
            core.add_obj(E.Null(ex.constant), Itr.empty())
        return core
    if isinstance(ex, io.QuadExpr):
        for (v1, v2), coef in ex.terms.items():
            m1, m2 = v1.index_type == "Measure", v2.index_type == "Measure"
            if m1 and m2:
                warnings.warn(_ObjMeasureExpansionWarn)
                core = _add_generic_objective_term(core, expand_measures(coef * v1 * v2, data), data)
            elif m1:
                core = _add_objective_aff_term(core, coef * v2, v1, data)
            else:
                core = _add_objective_aff_term(core, coef * v1, v2, data)
        return _add_objective(core, ex.aff, data, inf_model)
    if any(v.index_type == "Measure" for v in io.all_expression_variables(ex)):
        warnings.warn(_ObjMeasureExpansionWarn)
    return _add_generic_objective_term(core, expand_measures(ex, data), data)


def expand_measures(ex, data):
    """``InfiniteOpt.expand_measures``: every measure becomes the explicit weighted sum over its supports,
    with the measured parameter's variables replaced by point / semi-infinite variables."""
    def expand_one(mref: io.Measure):
        inner = expand_measures(mref.expr, data)
        supps, coeffs = _measure_data(mref, data)
        total = 0.0
        K = supps.shape[-1]
        for k in range(K):
            sval = {p: (float(supps[k]) if supps.ndim == 1 else float(supps[p.pos, k])) for p in mref.prefs}

            def at(v):
                if isinstance(v, io.InfiniteParameter) and v in sval:
                    return sval[v]
                if isinstance(v, (io.InfiniteVariable, io.Derivative)) and any(p in sval for p in v.prefs):
                    return v(*[sval.get(p, p) for p in v.prefs])
                if isinstance(v, io.SemiInfiniteVariable) and any(p in sval for p in v.prefs):
                    vals = [v.fixed[i] if i in v.fixed else sval.get(p, p) for i, p in enumerate(v.base.prefs)]
                    return v.base(*vals)
                if isinstance(v, io.ParameterFunction) and any(p in sval for p in v.prefs):
                    if all(p in sval for p in v.prefs):
                        return float(v.func(*[sval[p] for p in v.prefs]))
                    # partially evaluated: a semi-infinite parameter function (transform.jl:207-211 maps it to
                    # the parameter function's θ block with the fixed support's index; test/solve.jl:120)
                    return v.model._semi(v, {i: sval[p] for i, p in enumerate(v.prefs) if p in sval})
                return v
            total = total + float(coeffs[k]) * io.map_expression(at, inner)
        return total
    return io.map_expression(lambda v: expand_one(v) if v.index_type == "Measure" else v, ex) \
        if not io._is_num(ex) else ex


# ---- transform.jl:771-839 ------------------------------------------------------------------------------------------------
def build_exa_core(core: ExaCore, data: ExaMappingData, inf_model: io.InfiniteModel) -> ExaCore:
    """``build_exa_core!`` — the ORDER below defines x, θ, row and COO offsets (transform.jl:777-794)."""
    _build_base_iterators(data, inf_model)
    core = _add_finite_parameters(core, data, inf_model)
    core = _add_finite_variables(core, data, inf_model)
    core = _add_infinite_variables(core, data, inf_model)
    core = _add_parameter_functions(core, data, inf_model)
    _add_semi_infinite_variables(core, data, inf_model)
    _add_point_variables(core, data, inf_model)
    core = _add_constraints(core, data, inf_model)
    core = _add_derivative_approximations(core, data, inf_model)
    core = _add_collocation_restrictions(core, data, inf_model)
    if inf_model.objective_sense is not None:
        core = _add_objective(core, inf_model.objective_expr, data, inf_model)
    return core


def exa_core(inf_model: io.InfiniteModel, data: Optional[ExaMappingData] = None) -> Tuple[ExaCore, ExaMappingData]:
    """``ExaModels.ExaCore(inf_model, data; backend)`` (transform.jl:808-817)."""
    data = data or ExaMappingData()
    minimize = inf_model.objective_sense != "Max"
    core = ExaCore(minimize=minimize)
    return build_exa_core(core, data, inf_model), data

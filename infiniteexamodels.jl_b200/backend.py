"""``ExaTranscriptionBackend`` — host-side mirror of the part of ``src/infiniteopt_backend.jl`` that owns the
evaluation path: build (``build_transformation_backend!`` :150-157), the solver hand-off
(``JuMP.optimize!`` :259-271 → ``initial_solve`` / ``resolve``, ext/*.jl), the in-place updates that must not
rebuild the plan (``update_parameter_value`` :511-548, ``update_start_value`` :551-592, ``warmstart_backend``
:595-615) and value / dual extraction (``map_value`` :448-488, ``map_dual`` :490-508).

The MOI attribute plumbing, status translation and option diffing of that file are
solver bookkeeping with no arithmetic and are out of scope.  The solver itself (MadNLP / Ipopt) is a
Julia package that is absent here: ``solver`` is any callable ``solver(model, x0, y0, **options) -> Results``
that drives the NLPModels callbacks of ``model`` (tests use a scipy interior-point driver).
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass
from typing import Callable, Dict, Optional

import numpy as np

from . import infopt as io
from . import lib as _lib
from .model import ExaModel
from .transform import ExaMappingData, _evaluate_parameter_function, _support_values, exa_core


@dataclass
class Results:
    """what the reference reads from ``SolverCore.GenericExecutionStats`` (infiniteopt_backend.jl:408,444,600-601)"""
    solution: np.ndarray
    multipliers: np.ndarray
    objective: float
    status: str = "first_order"
    iter: int = 0


class ExaTranscriptionBackend:
    def __init__(self, solver: Optional[Callable] = None, device: int = 0, flags: int = _lib.IEXA_F_DEFAULT,
                 library=None, **options):
        self.solver, self.device, self.flags, self.library = solver, device, flags, library
        self.options: Dict = dict(options)
        self.inf_model = self.core = self.model = self.data = self.results = None

    # ---- build_transformation_backend! (:150-157) ---------------------------------------------------
    def build_transformation_backend(self, inf_model: io.InfiniteModel):
        self.empty()
        self.inf_model = inf_model
        self.core, self.data = exa_core(inf_model, ExaMappingData())
        self.model = ExaModel(self.core, device=self.device, flags=self.flags, library=self.library)
        return self

    def empty(self):
        """``Base.empty!(backend)`` (:134-143)"""
        self.core = self.model = self.data = self.results = None

    def transformation_backend_ready(self) -> bool:
        return self.model is not None

    # ---- optimize! (:259-271) ---------------------------------------------------------------------------
    def optimize(self, **options):
        if self.solver is None:
            raise RuntimeError("No solver attached to the ExaTranscriptionBackend.")   # :260
        if self.model is None:
            raise RuntimeError("build_transformation_backend must be called first")
        opts = dict(self.options); opts.update(options)
        self.results = self.solver(self.model, self.model.meta.x0.copy(), self.model.meta.y0.copy(), **opts)
        return self.results

    # ---- in-place updates (no plan rebuild) -----------------------------------------------------------
    def update_parameter_value(self, pref, value) -> bool:
        """finite parameter: a number; parameter function: a new callable evaluated at every support
        combination (:511-548).  Returns False when the backend must be rebuilt (unknown parameter)."""
        if self.model is None or pref not in self.data.param_mappings:
            return False
        block = self.data.param_mappings[pref]
        if isinstance(pref, io.FiniteParameter):
            pref.value = float(value)
            vals = np.array([float(value)])
        else:
            pref.func = value
            groups = pref.groups
            dims = tuple(self.data.base_itrs[g - 1].K for g in groups)
            vals = _evaluate_parameter_function(value, self.data, groups, dims)
        self.model.set_parameter(block, vals)
        th = self.core.theta_vec
        th[block.offset:block.offset + block.length] = np.asarray(vals).reshape(-1, order="F")
        return True

    def update_start_value(self, vref, value) -> bool:
        """(:551-592) scalar for finite / point variables, scalar or array for infinite variables"""
        if self.model is None:
            return False
        x0 = self.model.meta.x0
        if vref in self.data.finvar_mappings:
            x0[self.data.finvar_mappings[vref] - 1] = float(value)
        elif vref in self.data.infvar_mappings:
            blk = self.data.infvar_mappings[vref]
            if callable(value):      # a function of the supports, e.g. set_start_value(x, t -> 42) (test/solve.jl:225)
                value = _evaluate_parameter_function(lambda *s: float(value(*s)) if np.ndim(s[0]) == 0 else np.vectorize(value)(*s),
                                                     self.data, vref.groups, blk.size)
            x0[blk.offset:blk.offset + blk.length] = np.broadcast_to(np.asarray(value, dtype=np.float64), blk.size).reshape(-1, order="F")
        else:
            return False
        _lib.check(self.model.L, self.model.L.iexa_set_vector(self.model.h, 0, x0.ctypes.data))
        return True

    def warmstart_backend(self) -> bool:
        """copy the previous solution / multipliers into x0 / y0 (:595-603)"""
        if self.results is None or self.model is None:
            return False
        m = self.model
        m.meta.x0[:] = self.results.solution
        m.meta.y0[:] = self.results.multipliers
        _lib.check(m.L, m.L.iexa_set_vector(m.h, 0, m.meta.x0.ctypes.data))
        _lib.check(m.L, m.L.iexa_set_vector(m.h, 5, m.meta.y0.ctypes.data))
        return True

    # ---- result queries ---------------------------------------------------------------------------------
    def objective_value(self) -> float:
        return self.results.objective

    def map_value(self, vref):
        """(:448-488) finite / point variable -> scalar; infinite variable -> array shaped by its groups;
        parameter function -> its θ block"""
        sol = self.results.solution
        if vref in self.data.finvar_mappings:
            return float(sol[self.data.finvar_mappings[vref] - 1])
        if vref in self.data.infvar_mappings:
            blk = self.data.infvar_mappings[vref]
            return sol[blk.offset:blk.offset + blk.length].reshape(blk.size, order="F")
        if vref in self.data.param_mappings:
            blk = self.data.param_mappings[vref]
            return self.model.θ[blk.offset:blk.offset + blk.length].reshape(blk.size, order="F")
        raise KeyError(vref)

    def map_dual(self, constr):
        """(:490-508) multipliers of the rows of one constraint"""
        con = self.data.constraint_mappings[id(constr)]
        return self.results.multipliers[con.offset:con.offset + con.size]

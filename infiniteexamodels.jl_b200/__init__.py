"""B200-native evaluation engine behind the NLPModels callback API for the models that
InfiniteExaModels.jl's ``ExaTranscriptionBackend`` produces (host-side Python mirror).

Import as ``iexa_b200`` (the directory name contains a dot; ``iexa_b200.py`` at the repo root
loads this package under that name).
"""
from .expr import *  # noqa: F401,F403
from .expr import DataSource, Const, Null, Var, Par, IndexExpr, nl_op, OP
from .core import ExaCore, Itr, Variable, Parameter, Constraint
from .model import (ExaModel, NLPModelMeta, obj, grad_, cons_, jac_structure_, jac_coord_,
                    hess_structure_, hess_coord_, eval3_, jprod_, jtprod_, hprod_, get_x0, get_y0, jac_is_csr, jac_csr_rowptr_, device_bytes,
                    algorithmic_bytes, launches_per_call)
from . import lib

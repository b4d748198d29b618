"""ctypes binding of the C ABI (include/iexa.h).  This is the same binding a Julia `ccall` shim
makes; nothing here computes — every evaluation runs in libiexa_b200.so."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))

IEXA_MEM_HOST, IEXA_MEM_DEVICE, IEXA_MEM_HOST_SAME_X = 0, 1, 2
IEXA_F_DEFAULT, IEXA_F_NO_SPECIALISE, IEXA_F_NO_DEVICE = 0, 1, 2
IEXA_OPT_SLOT_ORDER, IEXA_OPT_STRICT_IEEE = 1, 2
CB_OBJ, CB_GRAD, CB_CONS, CB_JAC, CB_HESS, CB_JPROD, CB_JTPROD, CB_HPROD, CB_EVAL3 = range(9)


class IexaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"iexa error {code}: {msg}")
        self.code = code


class Meta(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("nvar", "ncon", "npar", "nobj_gen", "ncon_gen", "nnzj", "nnzh",
                                         "nnzg", "loc_ncon", "loc_nnzj", "loc_nnzh")] + \
               [(n, C.c_int32) for n in ("minimize", "rank", "world", "device", "n_kernels_specialised", "pad")]


class ColGen(C.Structure):
    _fields_ = [("kind", C.c_int32), ("src", C.c_int32), ("n", C.c_int64), ("a", C.c_double), ("b", C.c_double)]


class Segment(C.Structure):
    _fields_ = [("global_start", C.c_int64), ("local_start", C.c_int64), ("length", C.c_int64)]


# every symbol include/iexa.h declares (tests check the library exports all of them)
SYMBOLS = [
    "iexa_last_error", "iexa_version", "iexa_plan_create", "iexa_plan_destroy", "iexa_set_option", "iexa_add_var",
    "iexa_add_par", "iexa_patch_var", "iexa_itr_base", "iexa_itr_generated", "iexa_add_par_function", "iexa_debug_get_column", "iexa_itr_product", "iexa_add_con",
    "iexa_add_obj", "iexa_finalize", "iexa_get_meta", "iexa_get_vector", "iexa_set_vector",
    "iexa_set_par", "iexa_set_par_stream", "iexa_get_par", "iexa_jac_structure", "iexa_hess_structure", "iexa_obj",
    "iexa_grad", "iexa_cons", "iexa_jac_coord", "iexa_hess_coord", "iexa_eval3", "iexa_jprod", "iexa_jtprod",
    "iexa_hprod", "iexa_obj_device", "iexa_host_register", "iexa_host_unregister", "iexa_segments", "iexa_shared_vars", "iexa_shared_ranges", "iexa_host_x_bytes", "iexa_x_ranges", "iexa_algorithmic_bytes",
    "iexa_launches_per_call", "iexa_engine_note", "iexa_debug_codegen_source", "iexa_debug_set_class_mode", "iexa_debug_codegen_compile", "iexa_debug_codegen_source_of", "iexa_debug_codegen_compile_of", "iexa_debug_cache_stats",
    "iexa_halo_create", "iexa_halo_export", "iexa_halo_connect", "iexa_halo_set_sends", "iexa_halo_set_recvs", "iexa_halo_exchange",
    "iexa_halo_allreduce_small", "iexa_halo_status", "iexa_halo_destroy",
    "iexa_csr_create", "iexa_csr_create_keyed", "iexa_coo_locality", "iexa_jac_is_csr", "iexa_jac_csr_rowptr", "iexa_device_bytes", "iexa_csr_destroy", "iexa_csr_nnz", "iexa_csr_pattern", "iexa_csr_apply",
]

_vp, _i64, _i32, _dbl, _u32 = C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.c_uint32


def _declare(L):
    def sig(name, res, *args):
        if hasattr(L, name):
            f = getattr(L, name)
            f.restype, f.argtypes = res, list(args)

    sig("iexa_last_error", C.c_char_p)
    sig("iexa_version", _i32)
    sig("iexa_plan_create", _i32, C.POINTER(_vp), _i32)
    sig("iexa_plan_destroy", _i32, _vp)
    sig("iexa_set_option", _i32, _vp, _i32, _i64)
    sig("iexa_add_var", _i32, _vp, _i64, _vp, _vp, _vp, C.POINTER(_i64))
    sig("iexa_add_par", _i32, _vp, _i64, _vp, C.POINTER(_i64))
    sig("iexa_patch_var", _i32, _vp, _i32, _i64, _dbl)
    sig("iexa_itr_base", _i32, _vp, _i64, _i32, _vp, _i32, _vp, C.POINTER(_i32))
    sig("iexa_itr_generated", _i32, _vp, _i64, _i32, _vp, _i32, _vp, _vp, C.POINTER(_i32))
    sig("iexa_add_par_function", _i32, _vp, _vp, _i32, _vp, _i32, _i32, C.POINTER(_i64))
    sig("iexa_debug_get_column", _i32, _vp, _i32, _i32, _vp)
    sig("iexa_itr_product", _i32, _vp, _i32, _vp, C.POINTER(_i32))
    sig("iexa_add_con", _i32, _vp, _vp, _i32, _vp, _i32, _i32, _dbl, _dbl, C.POINTER(_i64))
    sig("iexa_add_obj", _i32, _vp, _vp, _i32, _vp, _i32, _i32)
    sig("iexa_finalize", _i32, _vp, _i32, _i32, _i32, _u32)
    sig("iexa_get_meta", _i32, _vp, C.POINTER(Meta))
    sig("iexa_get_vector", _i32, _vp, _i32, _vp)
    sig("iexa_set_vector", _i32, _vp, _i32, _vp)
    sig("iexa_set_par", _i32, _vp, _i64, _i64, _vp)
    sig("iexa_set_par_stream", _i32, _vp, _i64, _i64, _vp, _vp)
    sig("iexa_get_par", _i32, _vp, _i64, _i64, _vp)
    sig("iexa_jac_structure", _i32, _vp, _vp, _vp, _i32, _i32, _vp)
    sig("iexa_hess_structure", _i32, _vp, _vp, _vp, _i32, _i32, _vp)
    sig("iexa_obj", _i32, _vp, _vp, C.POINTER(_dbl), _i32, _vp)
    sig("iexa_obj_device", _i32, _vp, _vp, _vp, _vp)
    sig("iexa_grad", _i32, _vp, _vp, _vp, _i32, _vp)
    sig("iexa_cons", _i32, _vp, _vp, _vp, _i32, _vp)
    sig("iexa_jac_coord", _i32, _vp, _vp, _vp, _i32, _vp)
    sig("iexa_hess_coord", _i32, _vp, _vp, _vp, _dbl, _vp, _i32, _vp)
    sig("iexa_eval3", _i32, _vp, _vp, _vp, _dbl, _vp, _vp, _vp, _i32, _vp)
    sig("iexa_jprod", _i32, _vp, _vp, _vp, _vp, _i32, _vp)
    sig("iexa_jtprod", _i32, _vp, _vp, _vp, _vp, _i32, _vp)
    sig("iexa_hprod", _i32, _vp, _vp, _vp, _vp, _dbl, _vp, _i32, _vp)
    sig("iexa_host_register", _i32, _vp, _vp, _i64)
    sig("iexa_host_unregister", _i32, _vp, _vp)
    sig("iexa_segments", _i64, _vp, _i32, _vp, _i64)
    sig("iexa_shared_vars", _i64, _vp, _vp, _i64)
    sig("iexa_x_ranges", _i64, _vp, _vp, _i64)
    sig("iexa_shared_ranges", _i64, _vp, _vp, _i64)
    sig("iexa_host_x_bytes", _i64, _vp)
    sig("iexa_algorithmic_bytes", _i64, _vp, _i32)
    sig("iexa_launches_per_call", _i32, _vp, _i32)
    sig("iexa_engine_note", C.c_char_p, _vp)
    sig("iexa_debug_codegen_source", _i64, _vp, _vp, _i64)
    sig("iexa_debug_codegen_compile", _i32, _vp, C.POINTER(_i64))
    sig("iexa_debug_set_class_mode", _i32, _vp, _i32)
    sig("iexa_debug_codegen_source_of", _i64, _vp, _i32, _vp, _i64)
    sig("iexa_debug_codegen_compile_of", _i32, _vp, _i32, C.POINTER(_i64))
    sig("iexa_debug_cache_stats", _i32, C.POINTER(_i32), C.POINTER(_i32))
    sig("iexa_csr_create", _i32, C.POINTER(_vp), _i64, _i64, _i64, _vp, _vp, _i32, _i32, _i32)
    sig("iexa_csr_create_keyed", _i32, C.POINTER(_vp), _i64, _i64, _i64, _vp, _vp, _i32, _vp, _i32, _i32)
    sig("iexa_coo_locality", _i32, _vp, _i32, _vp, _i32, _vp)
    sig("iexa_jac_is_csr", _i32, _vp, _vp)
    sig("iexa_device_bytes", _i32, _vp, _vp)
    sig("iexa_jac_csr_rowptr", _i32, _vp, _vp, _i32, _i32, _vp)
    sig("iexa_csr_destroy", _i32, _vp)
    sig("iexa_csr_nnz", _i64, _vp)
    sig("iexa_csr_pattern", _i32, _vp, _vp, _vp, _i32)
    sig("iexa_csr_apply", _i32, _vp, _vp, _vp, _i32, _vp)
    sig("iexa_halo_create", _i32, C.POINTER(_vp), _i32, _i32, _i32)
    sig("iexa_halo_export", _i32, _vp, _vp, _vp, C.POINTER(_i64), _vp)
    sig("iexa_halo_connect", _i32, _vp, _i32, _vp, _i64, _vp)
    sig("iexa_halo_set_sends", _i32, _vp, _i32, _i64, _vp)
    sig("iexa_halo_set_recvs", _i32, _vp, _i32, _vp)
    sig("iexa_halo_exchange", _i32, _vp, _vp, _vp)
    sig("iexa_halo_allreduce_small", _i32, _vp, _vp, _i32, _vp)
    sig("iexa_halo_status", _i64, _vp)
    sig("iexa_halo_destroy", _i32, _vp)
    # test-only entry points of tests/hostcheck (absent from the product library)
    sig("hostcheck_eval", _i32, _vp, _i32, _vp, _vp, _dbl, _vp)
    sig("hostcheck_eval_local", _i32, _vp, _i32, _vp, _vp, _dbl, _vp)
    sig("hostcheck_residency_misses", _i64)
    sig("hostcheck_residency_bytes", _i32, _vp, _vp)
    sig("hostcheck_eval_groups", _i32, _vp, _i32, _vp, _vp, _dbl, _vp, C.POINTER(_i32))
    sig("hostcheck_set_class_mode", _i32, _vp, _i32)
    sig("hostcheck_structure", _i32, _vp, _i32, _vp, _vp)
    sig("hostcheck_scatter_info", _i32, _vp, _i32, _vp)
    sig("hostcheck_prod", _i32, _vp, _i32, _i32, _vp, _vp, _vp, _dbl, _vp, _vp)
    sig("hostcheck_gen_stats", _i32, _vp, _i32, _i32, _vp)
    sig("hostcheck_grad_stats", _i32, _vp, _vp)
    return L


_LIBS = {}


def load(path: str | None = None):
    """Load the product library (building it in-tree if missing).  No fallback of any kind: a
    missing/unbuildable CUDA extension raises."""
    if path is None:
        path = os.path.join(HERE, "libiexa_b200.so")
        if not os.path.exists(path):
            from . import build as _b
            _b.build()
    path = os.path.abspath(path)
    if path not in _LIBS:
        _LIBS[path] = _declare(C.CDLL(path))
    return _LIBS[path]


def check(L, rc):
    if rc != 0:
        raise IexaError(rc, L.iexa_last_error().decode(errors="replace"))

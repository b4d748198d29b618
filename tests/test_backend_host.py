"""Host logic of ``ExaTranscriptionBackend`` without a GPU (plan built with IEXA_F_NO_DEVICE): the reference's
"Start value updates" test (test/solve.jl:211-239) — in-place x0 updates for finite and infinite variables, scalar and
function-valued, and the "needs a rebuild" answer for variables the built backend does not know."""
import numpy as np

import iexa_b200 as ex
from iexa_b200 import infopt as io
from iexa_b200.backend import ExaTranscriptionBackend


def test_start_value_updates(hostcheck_lib):
    m = io.InfiniteModel()
    t = m.infinite_parameter(0, 1, num_supports=3)
    x = m.variable(t)
    z = m.variable(start=3.0)
    m.constraint(x + z, "==", 1)
    b = ExaTranscriptionBackend(None, flags=ex.lib.IEXA_F_NO_DEVICE, library=hostcheck_lib).build_transformation_backend(m)
    assert b.transformation_backend_ready()
    handle = b.model.h.value
    zi = b.data.finvar_mappings[z]                                      # transformation_variable(z).i
    assert b.model.meta.x0[zi - 1] == 3.0
    assert b.update_start_value(z, 10) and b.model.meta.x0[zi - 1] == 10            # solve.jl:219-221
    var = b.data.infvar_mappings[x]
    rng = slice(var.offset, var.offset + var.length)
    assert b.update_start_value(x, 20) and (b.model.meta.x0[rng] == 20).all()       # solve.jl:222-224
    assert b.update_start_value(x, lambda tt: 42) and (b.model.meta.x0[rng] == 42).all()   # solve.jl:225-227
    assert b.update_start_value(x, lambda tt: 1 + tt) and np.allclose(b.model.meta.x0[rng], [1.0, 1.5, 2.0])
    # the plan itself holds the new start (what a solver would be handed), and nothing was rebuilt
    x0 = np.zeros(b.model.meta.nvar)
    assert b.model.L.iexa_get_vector(b.model.h, 0, x0.ctypes.data) == 0
    assert np.array_equal(x0, b.model.meta.x0) and b.model.h.value == handle
    # variables the built backend does not know: the update reports that a rebuild is needed (solve.jl:229-238)
    w = m.variable()
    assert b.update_start_value(w, 5) is False
    q = m.variable(t)
    assert b.update_start_value(q, np.sin) is False

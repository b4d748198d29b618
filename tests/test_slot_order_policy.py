"""The COO slot ORDER inside a generator is DATA (iexa_set_option IEXA_OPT_SLOT_ORDER), honoured by the oracle and by the
product alike: ExaModels' own order cannot be pinned here (the package is absent — SURVEY App. A.2), so the day a real
dump disagrees with the default hypothesis only the policy value changes, not a kernel.  Under BOTH policies: structure
bit-exact and values within the north-star tolerance between the plan compiler (per-generator programs and fused groups,
host executor) and the oracle; the two policies give the same MATRICES (a permutation of the slots inside each support)."""
import ctypes as C

import numpy as np
import pytest

import iexa_b200 as ex
from iexa_b200 import models
from conftest import assert_close, eval_point

CASES = {
    "ode_5x5": lambda: models.ode_5x5(),
    "quadrotor_oc": lambda: models.quadrotor(9, "oc"),
    "pandemic": lambda: models.pandemic(7, 3),
    "farmer": lambda: models.farmer(11),
}


def _opf():
    from iexa_b200 import opf
    from iexa_b200.transform import exa_core
    return exa_core(opf.opf(None, num_supports=6))[0]


CASES["opf_case3"] = _opf


def _hc(L, m, fn, which, n, x, y=None, sig=1.0):
    out = np.zeros(max(n, 1))
    args = [m.h, which, x.ctypes.data, None if y is None else y.ctypes.data, sig, out.ctypes.data]
    if fn == "hostcheck_eval_groups":
        args.append(C.byref(C.c_int32()))
    assert getattr(L, fn)(*args) == 0
    return out[:n]


@pytest.mark.parametrize("order", [0, 1, 2])
@pytest.mark.parametrize("name", list(CASES))
def test_product_and_oracle_agree_under_each_policy(name, order, hostcheck_lib):
    from oracle.oracle import OracleModel
    L = hostcheck_lib
    core = CASES[name]()
    om = OracleModel(core, slot_order=order)
    m = ex.ExaModel(core, flags=ex.lib.IEXA_F_NO_DEVICE, library=L, slot_order=order)
    assert (m.meta.nnzj, m.meta.nnzh) == (om.nnzj, om.nnzh)
    x, y = eval_point(core)
    x = np.where(np.isfinite(x), x, 0.0)
    for which, (ro, co) in ((0, om.jac_structure()), (1, om.hess_structure())):
        r = np.zeros(max(len(ro), 1), dtype=np.int64); c = np.zeros_like(r)
        assert L.hostcheck_structure(m.h, which, r.ctypes.data, c.ctypes.data) == 0
        assert (r[:len(ro)] == ro).all() and (c[:len(co)] == co).all(), f"structure {which}"
    for fn in ("hostcheck_eval", "hostcheck_eval_groups"):
        assert_close(_hc(L, m, fn, 3, om.nnzj, x), om.jac_coord(x), "jac")
        assert_close(_hc(L, m, fn, 4, om.nnzh, x, y, 0.7), om.hess_coord(x, y, 0.7), "hess")
        assert_close(_hc(L, m, fn, 1, om.nvar, x), om.grad(x), "grad")


@pytest.mark.parametrize("name", list(CASES))
def test_policies_permute_slots_but_not_the_matrices(name):
    from oracle.oracle import OracleModel
    import scipy.sparse as sp
    core = CASES[name]()
    x, y = eval_point(core)
    x = np.where(np.isfinite(x), x, 0.0)
    mats = []
    structs = []
    for order in (0, 1, 2):
        om = OracleModel(core, slot_order=order)
        jr, jc = om.jac_structure(); hr, hc = om.hess_structure()
        J = sp.coo_matrix((om.jac_coord(x), (jr - 1, jc - 1)), shape=(om.ncon, om.nvar)).tocsr()
        H = sp.coo_matrix((om.hess_coord(x, y, 0.7), (hr - 1, hc - 1)), shape=(om.nvar, om.nvar)).tocsr()
        mats.append((J, H)); structs.append((jr, jc, hr, hc))
    for other in mats[1:]:
        for a, b in zip(mats[0], other):
            d = abs(a - b)
            assert d.max() <= 1e-12 * max(1.0, abs(a).max()) if d.nnz else True
    if name in ("quadrotor_oc", "opf_case3"):  # trees with binary nodes over several variables: the ORDER really differs
        assert any(not np.array_equal(u, v) for u, v in zip(structs[0], structs[1]))


@pytest.mark.parametrize("name", list(CASES))
def test_row_sorted_policy_makes_the_coo_array_a_csr_array(name, hostcheck_lib):
    """IEXA_SLOT_ORDER_JAC_ROW_SORTED: rows contiguous and columns strictly increasing inside each row — in the oracle (brute
    force over every support) and in the plan compiler (structural analysis) alike; the row pointers the C ABI hands out
    plus jac_structure's columns plus the UNTOUCHED jac_coord! values are the CSR matrix scipy builds from the COO triplets."""
    from oracle.oracle import OracleModel
    import scipy.sparse as sp
    L = hostcheck_lib
    core = CASES[name]()
    om = OracleModel(core, slot_order=2)
    assert om.L.orc_jac_is_csr(om.h) == 1
    m = ex.ExaModel(core, flags=ex.lib.IEXA_F_NO_DEVICE, library=L, slot_order=2)
    assert ex.jac_is_csr(m)
    jr, jc = om.jac_structure()
    assert (np.diff(jr) >= 0).all()
    same_row = np.diff(jr) == 0
    assert (np.diff(jc)[same_row] > 0).all()
    rp = np.zeros(om.ncon + 1, dtype=np.int64)
    ex.jac_csr_rowptr_(m, rp)
    rp32 = np.zeros(om.ncon + 1, dtype=np.int32)
    ex.jac_csr_rowptr_(m, rp32)
    assert (rp32 == rp).all()
    x, _ = eval_point(core)
    x = np.where(np.isfinite(x), x, 0.0)
    vals = om.jac_coord(x)
    ref = sp.coo_matrix((vals, (jr - 1, jc - 1)), shape=(om.ncon, om.nvar)).tocsr()
    ref.sort_indices()
    mine = sp.csr_matrix((vals, (jc - 1).astype(np.int64), rp), shape=(om.ncon, om.nvar))
    assert mine.has_sorted_indices or True
    assert (mine.indptr == ref.indptr).all() and (mine.indices == ref.indices).all()
    assert np.array_equal(mine.data, ref.data)      # nothing summed, nothing moved: bit-identical
    # the default policy is NOT CSR for the models whose trees meet the variables out of column order
    m0 = ex.ExaModel(core, flags=ex.lib.IEXA_F_NO_DEVICE, library=L)
    assert not ex.jac_is_csr(m0)
    with pytest.raises(Exception):
        ex.jac_csr_rowptr_(m0, rp)


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("name", ["quadrotor_oc", "pandemic", "farmer"])
def test_row_sorted_row_pointers_of_a_shard(name, world, hostcheck_lib):
    """world > 1: a rank's row pointers index ITS slice of the value array; mapped back through iexa_segments they are the
    global row pointers of the rows the rank owns, and the ranks' rows tile the model."""
    L = hostcheck_lib
    core = CASES[name]()
    g = ex.ExaModel(core, flags=ex.lib.IEXA_F_NO_DEVICE, library=L, slot_order=2)
    rp_g = ex.jac_csr_rowptr_(g, np.zeros(g.meta.ncon + 1, dtype=np.int64))
    seen_rows = np.zeros(g.meta.ncon, dtype=np.int64)
    for r in range(world):
        m = ex.ExaModel(core, flags=ex.lib.IEXA_F_NO_DEVICE, library=L, slot_order=2, rank=r, world=world)
        assert ex.jac_is_csr(m)
        rp = ex.jac_csr_rowptr_(m, np.zeros(m.loc_ncon + 1, dtype=np.int64))
        assert rp[0] == 0 and rp[-1] == m.loc_nnzj and (np.diff(rp) >= 0).all()

        def segs(which):
            n = L.iexa_segments(m.h, which, None, 0)
            sg = (ex.lib.Segment * max(n, 1))()
            L.iexa_segments(m.h, which, sg, n)
            return [(q.global_start, q.local_start, q.length) for q in sg[:n]]
        # local slot -> global slot
        l2g = np.full(max(m.loc_nnzj, 1), -1, dtype=np.int64)
        for gs, ls, ln in segs(1):
            l2g[ls:ls + ln] = np.arange(gs, gs + ln)
        for gs, ls, ln in segs(0):          # rows
            seen_rows[gs:gs + ln] += 1
            for i in range(ln):
                a, b = rp[ls + i], rp[ls + i + 1]
                assert b - a == rp_g[gs + i + 1] - rp_g[gs + i]
                if b > a:
                    assert l2g[a] == rp_g[gs + i] and l2g[b - 1] == rp_g[gs + i + 1] - 1
    assert (seen_rows == 1).all()


def test_row_sorted_policy_falls_back_per_generator_when_no_static_order_exists(hostcheck_lib):
    """A generator whose two variables swap their column order along the iterator keeps the default slot order; the model then
    reports 'not CSR' instead of a wrong pattern (oracle and plan compiler agree on that too)."""
    from oracle.oracle import OracleModel
    L = hostcheck_lib
    c = ex.ExaCore()
    ds = ex.DataSource()
    v = c.add_var(6)
    it = ex.Itr(3, {"a": np.array([1, 2, 5]), "b": np.array([4, 3, 2])})    # a < b, a < b, a > b
    c.add_con(v[ds.a] * 2.0 + v[ds.b] * v[ds.b], it, 0.0, 0.0)
    c.add_obj(v[1] * v[1], ex.Itr.empty())
    om = OracleModel(c, slot_order=2)
    m = ex.ExaModel(c, flags=ex.lib.IEXA_F_NO_DEVICE, library=L, slot_order=2)
    assert om.L.orc_jac_is_csr(om.h) == 0 and not ex.jac_is_csr(m)
    jr, jc = om.jac_structure()
    r = np.zeros(len(jr), dtype=np.int64); cc = np.zeros_like(r)
    assert L.hostcheck_structure(m.h, 0, r.ctypes.data, cc.ctypes.data) == 0
    assert (r == jr).all() and (cc == jc).all()


def test_options_are_frozen_once_a_generator_exists(hostcheck_lib):
    L = hostcheck_lib
    m = ex.ExaModel(models.farmer(3), flags=ex.lib.IEXA_F_NO_DEVICE, library=L)
    assert L.iexa_set_option(m.h, ex.lib.IEXA_OPT_SLOT_ORDER, 1) != 0
    h = C.c_void_p()
    assert L.iexa_plan_create(C.byref(h), 1) == 0
    assert L.iexa_set_option(h, ex.lib.IEXA_OPT_SLOT_ORDER, 7) != 0 and b"policy" in L.iexa_last_error()
    assert L.iexa_set_option(h, 99, 0) != 0
    assert L.iexa_set_option(h, ex.lib.IEXA_OPT_SLOT_ORDER, 1) == 0 and L.iexa_set_option(h, ex.lib.IEXA_OPT_STRICT_IEEE, 1) == 0
    L.iexa_plan_destroy(h)


@pytest.mark.parametrize("seed", range(0, 40, 3))
def test_row_sorted_policy_on_random_models(seed, hostcheck_lib):
    """fuzzed generators (arbitrary integer columns, product iterators, repeated variables): the plan compiler's structural
    decision and the oracle's brute-force decision agree generator by generator — including 'no static order' — and the structure /
    values agree under the policy; when every generator is sortable the COO pattern is CSR"""
    import test_fuzz
    from oracle.oracle import OracleModel
    L = hostcheck_lib
    core = test_fuzz.random_model(seed, K1=37, K2=3)[0]
    om = OracleModel(core, slot_order=2)
    m = ex.ExaModel(core, flags=ex.lib.IEXA_F_NO_DEVICE, library=L, slot_order=2)
    is_csr = ex.jac_is_csr(m)
    assert bool(om.L.orc_jac_is_csr(om.h)) == is_csr
    ro, co = om.jac_structure()
    r = np.zeros(max(len(ro), 1), dtype=np.int64); c = np.zeros_like(r)
    assert L.hostcheck_structure(m.h, 0, r.ctypes.data, c.ctypes.data) == 0
    assert (r[:len(ro)] == ro).all() and (c[:len(co)] == co).all()
    x, _ = eval_point(core)
    x = np.where(np.isfinite(x), x, 0.0)
    assert_close(_hc(L, m, "hostcheck_eval_groups", 3, om.nnzj, x), om.jac_coord(x), "jac")
    assert_close(_hc(L, m, "hostcheck_eval", 3, om.nnzj, x), om.jac_coord(x), "jac")
    if is_csr and len(ro) > 1:
        same = np.diff(ro) == 0
        assert (np.diff(ro) >= 0).all() and (np.diff(co)[same] > 0).all()

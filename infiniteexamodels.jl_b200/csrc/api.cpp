// api.cpp — extern "C" entry points declared in include/iexa.h.
#include <algorithm>
#include <cstring>
#include <memory>
#include <string>

#include "../../include/iexa.h"
#include "engine.hpp"
#include "plan.hpp"

#include "api_internal.hpp"
#include "codegen.hpp"

namespace iexa { thread_local std::string g_last_error; }
#define g_err iexa::g_last_error
static int32_t fail(int32_t code, const std::string &msg) { g_err = msg; return code; }

#define GUARD_BEGIN try {
#define GUARD_END                                                              \
  }                                                                            \
  catch (const std::invalid_argument &e) { return fail(IEXA_ERR_INVALID, e.what()); } \
  catch (const std::logic_error &e) { return fail(IEXA_ERR_STATE, e.what()); } \
  catch (const std::bad_alloc &) { return fail(IEXA_ERR_NOMEM, "out of host memory"); } \
  catch (const std::exception &e) { return fail(IEXA_ERR_INVALID, e.what()); } \
  catch (...) { return fail(IEXA_ERR_INVALID, "unknown C++ exception"); }

#define NEED_PLAN(p) if (!(p)) return fail(IEXA_ERR_INVALID, "null plan")
#define NEED_ENGINE(p)                                                                          \
  NEED_PLAN(p);                                                                                 \
  if (!(p)->plan.finalized) return fail(IEXA_ERR_STATE, "plan not finalized");                  \
  if (!(p)->engine) return fail(IEXA_ERR_CUDA, "plan has no CUDA engine (finalized with IEXA_F_NO_DEVICE or no GPU): there is no CPU fallback")

extern "C" {

const char *iexa_last_error(void) { return g_err.c_str(); }
int32_t iexa_version(void) { return IEXA_VERSION; }

int32_t iexa_plan_create(iexa_plan **out, int32_t minimize) {
  GUARD_BEGIN
  if (!out) return fail(IEXA_ERR_INVALID, "null out");
  *out = new iexa_plan();
  (*out)->plan.minimize = minimize != 0;
  return IEXA_OK;
  GUARD_END
}
int32_t iexa_plan_destroy(iexa_plan *p) {
  GUARD_BEGIN
  delete p;
  return IEXA_OK;
  GUARD_END
}

int32_t iexa_set_option(iexa_plan *p, int32_t key, int64_t value) {
  GUARD_BEGIN
  NEED_PLAN(p);
  if (p->plan.finalized || !p->plan.objs.empty() || !p->plan.cons.empty())
    return fail(IEXA_ERR_STATE, "options must be set before the first generator is added");
  switch (key) {
    case IEXA_OPT_SLOT_ORDER:
      if (value != IEXA_SLOT_ORDER_LEFT_TO_RIGHT && value != IEXA_SLOT_ORDER_RIGHT_TO_LEFT && value != IEXA_SLOT_ORDER_JAC_ROW_SORTED) return fail(IEXA_ERR_INVALID, "unknown slot-order policy");
      p->plan.opt_slot_order = (int)value;
      return IEXA_OK;
    case IEXA_OPT_STRICT_IEEE:
      p->plan.opt_strict = value != 0;
      return IEXA_OK;
    default: return fail(IEXA_ERR_INVALID, "unknown option");
  }
  GUARD_END
}

int32_t iexa_add_var(iexa_plan *p, int64_t n, const double *x0, const double *lvar, const double *uvar,
                     int64_t *offset_out) {
  GUARD_BEGIN
  NEED_PLAN(p);
  if (p->plan.finalized) return fail(IEXA_ERR_STATE, "plan already finalized");
  if (n < 0) return fail(IEXA_ERR_INVALID, "negative length");
  int64_t off = p->plan.add_var(n, x0, lvar, uvar);
  if (offset_out) *offset_out = off;
  return IEXA_OK;
  GUARD_END
}
int32_t iexa_add_par(iexa_plan *p, int64_t n, const double *vals, int64_t *offset_out) {
  GUARD_BEGIN
  NEED_PLAN(p);
  if (p->plan.finalized) return fail(IEXA_ERR_STATE, "plan already finalized");
  if (n < 0 || (n > 0 && !vals)) return fail(IEXA_ERR_INVALID, "bad parameter block");
  int64_t off = p->plan.add_par(n, vals);
  if (offset_out) *offset_out = off;
  return IEXA_OK;
  GUARD_END
}
int32_t iexa_patch_var(iexa_plan *p, int32_t which, int64_t i, double value) {
  GUARD_BEGIN
  NEED_PLAN(p);
  if (p->plan.finalized && which != 0) return fail(IEXA_ERR_STATE, "bounds are frozen after finalize");
  if (i < 1 || i > p->plan.nvar) return fail(IEXA_ERR_INVALID, "variable index out of range");
  if (which >= 0 && which <= 2) p->plan.materialise_vars();
  std::vector<double> *v = which == 0 ? &p->plan.x0 : which == 1 ? &p->plan.lvar : which == 2 ? &p->plan.uvar : nullptr;
  if (!v) return fail(IEXA_ERR_INVALID, "which must be 0,1,2");
  (*v)[i - 1] = value;
  return IEXA_OK;
  GUARD_END
}

int32_t iexa_itr_base(iexa_plan *p, int64_t K, int32_t n_int, const int64_t *const *int_cols, int32_t n_fp,
                      const double *const *fp_cols, int32_t *itr_out) {
  GUARD_BEGIN
  NEED_PLAN(p);
  if (p->plan.finalized) return fail(IEXA_ERR_STATE, "plan already finalized");
  if (n_int < 0 || n_fp < 0 || !itr_out) return fail(IEXA_ERR_INVALID, "bad iterator arguments");
  *itr_out = p->plan.itr_base(K, n_int, int_cols, n_fp, fp_cols);
  return IEXA_OK;
  GUARD_END
}
int32_t iexa_itr_generated(iexa_plan *p, int64_t K, int32_t n_int, const int64_t *const *int_cols, int32_t n_fp,
                           const iexa_colgen *gens, const double *const *fp_cols, int32_t *itr_out) {
  GUARD_BEGIN
  NEED_PLAN(p);
  if (p->plan.finalized) return fail(IEXA_ERR_STATE, "plan already finalized");
  if (n_int < 0 || n_fp < 0 || !itr_out || (n_fp > 0 && !gens)) return fail(IEXA_ERR_INVALID, "bad iterator arguments");
  *itr_out = p->plan.itr_generated(K, n_int, int_cols, n_fp, gens, fp_cols);
  return IEXA_OK;
  GUARD_END
}
int32_t iexa_add_par_function(iexa_plan *p, const iexa_node *nodes, int32_t n_nodes, const iexa_index *idx, int32_t n_idx,
                              int32_t itr, int64_t *offset_out) {
  GUARD_BEGIN
  NEED_PLAN(p);
  if (!nodes || n_idx < 0 || (n_idx > 0 && !idx)) return fail(IEXA_ERR_INVALID, "bad tape arguments");
  int64_t off = p->plan.add_par_function(nodes, n_nodes, idx, n_idx, itr);
  if (offset_out) *offset_out = off;
  return IEXA_OK;
  GUARD_END
}
int32_t iexa_debug_get_column(iexa_plan *p, int32_t itr, int32_t col, double *out_host) {
  GUARD_BEGIN
  NEED_PLAN(p);
  if (itr < 0 || itr >= (int32_t)p->plan.itrs.size() || !out_host) return fail(IEXA_ERR_INVALID, "bad iterator id");
  const iexa::Iterator &it = p->plan.itrs[itr];
  if (col < 0 || col >= (int32_t)it.fp_cols.size()) return fail(IEXA_ERR_INVALID, "bad fp column");
  const iexa::ColRef &r = it.fp_cols[col];
  const iexa::HostColumn &hc = p->plan.columns[r.col];
  if (p->engine) {
    std::string err;
    int rc = p->engine->get_column(r.col, out_host, err);
    if (rc) return fail(rc, err);
    return IEXA_OK;
  }
  for (int64_t j = 0; j < hc.K; ++j) out_host[j] = hc.fp(j, p->plan.columns); // host-only plan: the same closed forms
  return IEXA_OK;
  GUARD_END
}

int32_t iexa_itr_product(iexa_plan *p, int32_t n, const int32_t *itrs, int32_t *itr_out) {
  GUARD_BEGIN
  NEED_PLAN(p);
  if (p->plan.finalized) return fail(IEXA_ERR_STATE, "plan already finalized");
  if (n < 0 || !itr_out) return fail(IEXA_ERR_INVALID, "bad product arguments");
  *itr_out = p->plan.itr_product(n, itrs);
  return IEXA_OK;
  GUARD_END
}

int32_t iexa_add_con(iexa_plan *p, const iexa_node *nodes, int32_t n_nodes, const iexa_index *idx, int32_t n_idx,
                     int32_t itr, double lcon, double ucon, int64_t *row_offset_out) {
  GUARD_BEGIN
  NEED_PLAN(p);
  if (!nodes || n_idx < 0 || (n_idx > 0 && !idx)) return fail(IEXA_ERR_INVALID, "bad tape arguments");
  int64_t off = p->plan.add_con(nodes, n_nodes, idx, n_idx, itr, lcon, ucon);
  if (row_offset_out) *row_offset_out = off;
  return IEXA_OK;
  GUARD_END
}
int32_t iexa_add_obj(iexa_plan *p, const iexa_node *nodes, int32_t n_nodes, const iexa_index *idx, int32_t n_idx,
                     int32_t itr) {
  GUARD_BEGIN
  NEED_PLAN(p);
  if (!nodes || n_idx < 0 || (n_idx > 0 && !idx)) return fail(IEXA_ERR_INVALID, "bad tape arguments");
  p->plan.add_obj(nodes, n_nodes, idx, n_idx, itr);
  return IEXA_OK;
  GUARD_END
}

int32_t iexa_finalize(iexa_plan *p, int32_t device, int32_t rank, int32_t world, uint32_t flags) {
  GUARD_BEGIN
  NEED_PLAN(p);
  if (p->plan.finalized) return fail(IEXA_ERR_STATE, "plan already finalized");
  p->plan.layout(rank, world);
  p->plan.device = device;
  p->flags = flags;
  if (flags & IEXA_F_NO_DEVICE) return IEXA_OK;
  std::string err;
  iexa::Engine *e = iexa::make_cuda_engine(p->plan, device, flags, err);
  if (!e) return fail(IEXA_ERR_CUDA, err);
  p->engine.reset(e);
  return IEXA_OK;
  GUARD_END
}

int32_t iexa_get_meta(const iexa_plan *p, iexa_meta *m) {
  GUARD_BEGIN
  NEED_PLAN(p);
  if (!m) return fail(IEXA_ERR_INVALID, "null out");
  const iexa::Plan &P = p->plan;
  std::memset(m, 0, sizeof *m);
  m->nvar = P.nvar; m->ncon = P.ncon; m->npar = P.npar;
  m->nobj_gen = (int64_t)P.objs.size(); m->ncon_gen = (int64_t)P.cons.size();
  m->nnzj = P.nnzj; m->nnzh = P.nnzh; m->nnzg = P.nnzg;
  m->loc_ncon = P.loc_ncon; m->loc_nnzj = P.loc_nnzj; m->loc_nnzh = P.loc_nnzh;
  m->minimize = P.minimize; m->rank = P.rank; m->world = P.world; m->device = P.device;
  m->n_kernels_specialised = p->engine ? p->engine->n_specialised() : 0;
  return IEXA_OK;
  GUARD_END
}

int32_t iexa_get_vector(const iexa_plan *p, int32_t which, double *out) {
  GUARD_BEGIN
  NEED_PLAN(p);
  const iexa::Plan &P = p->plan;
  if (!out || which < 0 || which > 5) return fail(IEXA_ERR_INVALID, "bad vector selector");
  if (which == 3 || which == 4) { P.fill_con_bounds(which == 4, out); return IEXA_OK; }
  if (which == 5) {
    if (P.y0.empty()) std::fill(out, out + P.ncon, 0.0);
    else std::memcpy(out, P.y0.data(), P.y0.size() * 8);
    return IEXA_OK;
  }
  const_cast<iexa::Plan &>(P).materialise_vars();
  const std::vector<double> *v = which == 0 ? &P.x0 : which == 1 ? &P.lvar : &P.uvar;
  if (!v->empty()) std::memcpy(out, v->data(), v->size() * 8);
  return IEXA_OK;
  GUARD_END
}
int32_t iexa_set_vector(iexa_plan *p, int32_t which, const double *in) {
  GUARD_BEGIN
  NEED_PLAN(p);
  iexa::Plan &P = p->plan;
  if (!in || (which != 0 && which != 5)) return fail(IEXA_ERR_INVALID, "only x0 (0) and y0 (5) can be set");
  if (which == 5) { P.y0.assign(in, in + P.ncon); return IEXA_OK; }
  P.materialise_vars();
  if (!P.x0.empty()) std::memcpy(P.x0.data(), in, P.x0.size() * 8);
  return IEXA_OK;
  GUARD_END
}

static int32_t set_par_impl(iexa_plan *p, int64_t off, int64_t n, const double *vals, void *stream, bool device_sync) {
  GUARD_BEGIN
  NEED_PLAN(p);
  if (off < 0 || n < 0 || off + n > p->plan.npar || (n > 0 && !vals)) return fail(IEXA_ERR_INVALID, "parameter range out of bounds");
  std::memcpy(p->plan.theta.data() + off, vals, (size_t)n * 8);
  if (p->engine) {
    std::string err;
    int rc = p->engine->set_par(off, n, vals, stream, device_sync, err);
    if (rc) return fail(rc, err);
  }
  return IEXA_OK;
  GUARD_END
}
int32_t iexa_set_par(iexa_plan *p, int64_t off, int64_t n, const double *vals) { return set_par_impl(p, off, n, vals, nullptr, true); }
int32_t iexa_set_par_stream(iexa_plan *p, int64_t off, int64_t n, const double *vals, void *stream) {
  return set_par_impl(p, off, n, vals, stream, false);
}
int32_t iexa_get_par(const iexa_plan *p, int64_t off, int64_t n, double *vals) {
  GUARD_BEGIN
  NEED_PLAN(p);
  if (off < 0 || n < 0 || off + n > p->plan.npar || (n > 0 && !vals)) return fail(IEXA_ERR_INVALID, "parameter range out of bounds");
  std::memcpy(vals, p->plan.theta.data() + off, (size_t)n * 8);
  return IEXA_OK;
  GUARD_END
}

#define ENGINE_CALL(expr)          \
  GUARD_BEGIN                      \
  NEED_ENGINE(p);                  \
  std::string err;                 \
  int rc = p->engine->expr;        \
  if (rc) return fail(rc, err);    \
  return IEXA_OK;                  \
  GUARD_END

int32_t iexa_jac_structure(iexa_plan *p, void *rows, void *cols, int32_t idx_bytes, int32_t memspace, void *stream) {
  ENGINE_CALL(structure(0, rows, cols, idx_bytes, memspace, stream, err))
}
int32_t iexa_hess_structure(iexa_plan *p, void *rows, void *cols, int32_t idx_bytes, int32_t memspace, void *stream) {
  ENGINE_CALL(structure(1, rows, cols, idx_bytes, memspace, stream, err))
}
int32_t iexa_device_bytes(const iexa_plan *p, int64_t *out6) {
  GUARD_BEGIN
  NEED_PLAN(p);
  if (!out6) return fail(IEXA_ERR_INVALID, "null output");
  if (!p->plan.finalized || !p->engine) return fail(IEXA_ERR_STATE, "no device engine");
  p->engine->device_bytes(out6);
  return IEXA_OK;
  GUARD_END
}
int32_t iexa_jac_is_csr(const iexa_plan *p, int32_t *is_csr_out) {
  GUARD_BEGIN
  NEED_PLAN(p);
  if (!is_csr_out) return fail(IEXA_ERR_INVALID, "null output");
  if (!p->plan.finalized) return fail(IEXA_ERR_STATE, "plan not finalized");
  *is_csr_out = p->plan.jac_is_csr() ? 1 : 0;
  return IEXA_OK;
  GUARD_END
}
int32_t iexa_jac_csr_rowptr(iexa_plan *p, void *rowptr, int32_t idx_bytes, int32_t memspace, void *stream) {
  if (p && p->plan.finalized && memspace == IEXA_MEM_HOST && rowptr && (idx_bytes == 4 || idx_bytes == 8)) {
    // pure structure into a host buffer: answered from the plan (also on a build without a device)
    GUARD_BEGIN
    if (!p->plan.jac_is_csr()) return fail(IEXA_ERR_STATE, "the Jacobian COO order is not CSR (needs IEXA_SLOT_ORDER_JAC_ROW_SORTED and a static column order in every generator)");
    const std::vector<int64_t> rp = p->plan.jac_rowptr();
    if (idx_bytes == 8) std::memcpy(rowptr, rp.data(), rp.size() * 8);
    else {
      if (rp.back() > INT32_MAX) return fail(IEXA_ERR_INVALID, "nnzj does not fit Int32 row pointers");
      for (size_t i = 0; i < rp.size(); ++i) static_cast<int32_t *>(rowptr)[i] = (int32_t)rp[i];
    }
    return IEXA_OK;
    GUARD_END
  }
  ENGINE_CALL(jac_rowptr(rowptr, idx_bytes, memspace, stream, err))
}
int32_t iexa_coo_locality(iexa_plan *p, int32_t which, int32_t *keys, int32_t memspace, void *stream) {
  if (which != 0 && which != 1) return fail(IEXA_ERR_INVALID, "which must be 0 (Jacobian) or 1 (Hessian)");
  ENGINE_CALL(structure(2 + which, keys, nullptr, 4, memspace, stream, err))
}
int32_t iexa_obj(iexa_plan *p, const double *x, double *f_host, int32_t memspace, void *stream) {
  ENGINE_CALL(obj(x, f_host, memspace, stream, err))
}
int32_t iexa_obj_device(iexa_plan *p, const double *x_dev, double *f_dev, void *stream) {
  ENGINE_CALL(obj_device(x_dev, f_dev, stream, err))
}
int32_t iexa_grad(iexa_plan *p, const double *x, double *g, int32_t memspace, void *stream) {
  ENGINE_CALL(grad(x, g, memspace, stream, err))
}
int32_t iexa_cons(iexa_plan *p, const double *x, double *c, int32_t memspace, void *stream) {
  ENGINE_CALL(cons(x, c, memspace, stream, err))
}
int32_t iexa_jac_coord(iexa_plan *p, const double *x, double *vals, int32_t memspace, void *stream) {
  ENGINE_CALL(jac(x, vals, memspace, stream, err))
}
int32_t iexa_hess_coord(iexa_plan *p, const double *x, const double *y, double obj_weight, double *vals,
                        int32_t memspace, void *stream) {
  ENGINE_CALL(hess(x, y, obj_weight, vals, memspace, stream, err))
}
int32_t iexa_eval3(iexa_plan *p, const double *x, const double *y, double obj_weight, double *c, double *jac_vals, double *hess_vals,
                   int32_t memspace, void *stream) {
  ENGINE_CALL(eval3(x, y, obj_weight, c, jac_vals, hess_vals, memspace, stream, err))
}
int32_t iexa_jprod(iexa_plan *p, const double *x, const double *v, double *Jv, int32_t memspace, void *stream) {
  ENGINE_CALL(jprod(x, v, Jv, memspace, stream, err))
}
int32_t iexa_jtprod(iexa_plan *p, const double *x, const double *v, double *Jtv, int32_t memspace, void *stream) {
  ENGINE_CALL(jtprod(x, v, Jtv, memspace, stream, err))
}
int32_t iexa_hprod(iexa_plan *p, const double *x, const double *y, const double *v, double obj_weight, double *Hv,
                   int32_t memspace, void *stream) {
  ENGINE_CALL(hprod(x, y, v, obj_weight, Hv, memspace, stream, err))
}

int32_t iexa_host_register(iexa_plan *p, void *buf, int64_t bytes) {
  ENGINE_CALL(host_register(buf, (size_t)bytes, err))
}
int32_t iexa_host_unregister(iexa_plan *p, void *buf) {
  ENGINE_CALL(host_unregister(buf, err))
}

// ---- sharding queries ----------------------------------------------------------------------------
int64_t iexa_segments(const iexa_plan *p, int32_t which, iexa_segment *out, int64_t cap) {
  if (!p || !p->plan.finalized) return -1;
  const iexa::Plan &P = p->plan;
  int64_t n = 0;
  auto push = [&](int64_t gs, int64_t ls, int64_t len) {
    if (len <= 0) return;
    if (out && n < cap) out[n] = iexa_segment{gs, ls, len};
    ++n;
  };
  if (which == 2)
    for (auto &g : P.objs) push(g.o2 + g.k0 * g.c.o2step, g.l2, (g.k1 - g.k0) * g.c.o2step);
  for (auto &g : P.cons) {
    if (which == 0) push(g.o0 + g.k0, g.l0, g.k1 - g.k0);
    else if (which == 1) push(g.o1 + g.k0 * g.c.o1step, g.l1, (g.k1 - g.k0) * g.c.o1step);
    else push(g.o2 + g.k0 * g.c.o2step, g.l2, (g.k1 - g.k0) * g.c.o2step);
  }
  return n;
}

// Gradient entries that receive contributions from MORE THAN ONE rank: for every rank r the cover U_r of the variable
// indices its share [k0_r, k1_r) of every objective generator reaches through any first-order slot (Plan::index_range —
// shifted references y[i-1] at the shard boundaries, (k / div) % mod indices of product or restricted iterators and
// constant indices are all index ranges); shared = union over r < r' of U_r ∩ U_r'.  Every rank computes the same set
// from the plan alone.  The cover is conservative: an entry may be listed although a single rank contributes to it
// (the others then add zero), never the other way round.
static std::vector<std::pair<int64_t, int64_t>> shared_ranges(const iexa::Plan &P) { // [lo, hi] 1-based inclusive, merged
  typedef std::pair<int64_t, int64_t> Iv;
  auto merge = [](std::vector<Iv> &v) {
    std::sort(v.begin(), v.end());
    std::vector<Iv> m;
    for (auto &r : v) {
      if (!m.empty() && r.first <= m.back().second + 1) m.back().second = std::max(m.back().second, r.second);
      else m.push_back(r);
    }
    v.swap(m);
  };
  std::vector<std::vector<Iv>> U(P.world);
  for (int r = 0; r < P.world; ++r) {
    for (auto &g : P.objs) {
      int64_t k0, k1;
      iexa::Plan::shard_range(g.K, r, P.world, k0, k1);
      if (k1 <= k0) continue;
      const iexa::Iterator &it = P.itrs[g.itr];
      for (int32_t s : g.c.jac_slot) {
        int64_t lo, hi;
        P.index_range(it, g.c.int_cols, g.c.uidx[s], k0, k1, lo, hi);
        lo = std::max<int64_t>(lo, 1); hi = std::min<int64_t>(hi, P.nvar);
        if (lo <= hi) U[r].push_back({lo, hi});
      }
    }
    merge(U[r]);
  }
  std::vector<Iv> sh;
  for (int a = 0; a < P.world; ++a)
    for (int b = a + 1; b < P.world; ++b) {
      size_t i = 0, j = 0;
      while (i < U[a].size() && j < U[b].size()) {
        const int64_t lo = std::max(U[a][i].first, U[b][j].first), hi = std::min(U[a][i].second, U[b][j].second);
        if (lo <= hi) sh.push_back({lo, hi});
        if (U[a][i].second < U[b][j].second) ++i; else ++j;
      }
    }
  merge(sh);
  return sh;
}

int64_t iexa_shared_ranges(const iexa_plan *p, iexa_segment *out, int64_t cap) {
  if (!p || !p->plan.finalized) return -1;
  auto sh = shared_ranges(p->plan);
  for (size_t i = 0; i < sh.size() && (int64_t)i < cap && out; ++i) out[i] = iexa_segment{sh[i].first - 1, sh[i].first - 1, sh[i].second - sh[i].first + 1};
  return (int64_t)sh.size();
}

int64_t iexa_shared_vars(const iexa_plan *p, int64_t *out, int64_t cap) {
  if (!p || !p->plan.finalized) return -1;
  auto sh = shared_ranges(p->plan);
  int64_t n = 0;
  for (auto &r : sh)
    for (int64_t i = r.first; i <= r.second; ++i) { if (out && n < cap) out[n] = i; ++n; }
  return n;
}

int64_t iexa_x_ranges(const iexa_plan *p, iexa_segment *out, int64_t cap) {
  if (!p || !p->plan.finalized) return -1;
  const iexa::Plan &P = p->plan;
  std::vector<std::pair<int64_t, int64_t>> m = P.x_read_ranges();
  for (size_t i = 0; i < m.size() && (int64_t)i < cap && out; ++i) out[i] = iexa_segment{m[i].first - 1, m[i].first - 1, m[i].second - m[i].first + 1};
  return (int64_t)m.size();
}

int64_t iexa_host_x_bytes(const iexa_plan *p) {
  if (!p || !p->plan.finalized) return -1;
  const iexa::Plan &P = p->plan;
  if (P.world == 1) return 8 * P.nvar;
  int64_t n = 0;
  for (auto &r : P.x_read_ranges()) n += r.second - r.first + 1;
  return 8 * n;
}

// ---- algorithmic bytes --------------------------------------------------------
int64_t iexa_algorithmic_bytes(const iexa_plan *pc, int32_t which) {
  if (!pc || !pc->plan.finalized || which < 0 || which > 8) return -1;
  iexa_plan *p = const_cast<iexa_plan *>(pc);
  if (p->bytes_cache[which] >= 0) return p->bytes_cache[which];
  const iexa::Plan &P = p->plan;
  // (generator, program) pairs of the call; distinct inputs are counted once over ALL of them
  std::vector<std::pair<const iexa::Generator *, const iexa::Program *>> work;
  auto prog_of = [&](const iexa::Generator &g, int w) -> const iexa::Program & {
    return (w == 0 || w == 2) ? g.c.val : (w == 1 || w == 3) ? g.c.d1 : w == 4 ? g.c.d2 : w == 5 ? g.c.jv : w == 6 ? g.c.jtv : g.c.hv;
  };
  if (which == 8) { // iexa_eval3: value + first + second order of the constraints and second order of the objectives, inputs ONCE
    for (auto &g : P.objs) work.push_back({&g, &g.c.d2});
    for (auto &g : P.cons) { work.push_back({&g, &g.c.val}); work.push_back({&g, &g.c.d1}); work.push_back({&g, &g.c.d2}); }
  } else {
    const bool use_obj = which == 0 || which == 1 || which == 4 || which == 7;
    const bool use_con = which >= 2;
    if (use_obj) for (auto &g : P.objs) work.push_back({&g, &prog_of(g, which)});
    if (use_con) for (auto &g : P.cons) work.push_back({&g, &prog_of(g, which)});
  }
  std::vector<bool> xs((size_t)P.nvar, false), ts((size_t)P.npar, false), vs(which >= 5 && which <= 7 ? (size_t)P.nvar : 0, false);
  std::vector<std::vector<bool>> colseen(P.columns.size());
  std::vector<bool> yrow(which == 8 ? (size_t)P.ncon : 0, false);
  int64_t bytes = 0;
  for (auto &wk : work) {
    const iexa::Generator &g = *wk.first;
    const iexa::Program &pr = *wk.second;
    if (pr.nout == 0) continue;
    const iexa::Iterator &it = P.itrs[g.itr];
    std::vector<int32_t> lx, lp, lv;
    std::vector<int32_t> fcols, icols_used;
    std::vector<bool> iused(g.c.int_cols.size(), false);
    auto mark_idx = [&](int32_t islot) { for (auto &t : g.c.uidx[islot].terms) iused[t.first] = true; };
    for (const iexa::Instr &I : pr.code) {
      if (I.op == iexa::D_LOADX) { lx.push_back(I.a); mark_idx(I.a); }
      else if (I.op == iexa::D_LOADP) { lp.push_back(I.a); mark_idx(I.a); }
      else if (I.op == iexa::D_LOADV) { lv.push_back(I.a); mark_idx(I.a); }
      else if (I.op == iexa::D_SELNE) { mark_idx(I.a); mark_idx(I.b); }
      else if (I.op == iexa::D_FIELD) fcols.push_back(I.a);
      else if (I.op == iexa::D_SEL2) { mark_idx(I.a); mark_idx(I.b); }
    }
    if (which == 1) for (int32_t s : g.c.jac_slot) mark_idx(s);
    if (which == 6) for (int32_t s : g.c.jtv_slot) mark_idx(s);
    if (which == 7) for (int32_t s : g.c.hv_slot) mark_idx(s);
    for (int64_t k = g.k0; k < g.k1; ++k) {
      for (int32_t s : lx) { int64_t i = P.index_value(g, s, k) - 1; if (i >= 0 && i < P.nvar && !xs[i]) { xs[i] = true; bytes += 8; } }
      for (int32_t s : lp) { int64_t i = P.index_value(g, s, k) - 1; if (i >= 0 && i < P.npar && !ts[i]) { ts[i] = true; bytes += 8; } }
      for (int32_t s : lv) { int64_t i = P.index_value(g, s, k) - 1; if (i >= 0 && i < P.nvar && !vs[i]) { vs[i] = true; bytes += 8; } }
    }
    auto touch_col = [&](const iexa::ColRef &r, int esz) {
      const iexa::HostColumn &c = P.columns[r.col];
      if (c.is_int && c.affine) return; // index arithmetic, nothing is loaded
      auto &seen = colseen[r.col];
      if (seen.empty()) seen.assign((size_t)c.K, false);
      // entries reached by the local support range
      if (r.div == 1 && r.mod == g.K) {
        for (int64_t k = g.k0; k < g.k1; ++k) if (!seen[k]) { seen[k] = true; bytes += esz; }
      } else {
        for (int64_t k = g.k0; k < g.k1; ++k) { int64_t j = (k / r.div) % r.mod; if (!seen[j]) { seen[j] = true; bytes += esz; } }
      }
    };
    for (int32_t s : fcols) touch_col(it.fp_cols[g.c.fp_cols[s]], 8);
    for (size_t s = 0; s < iused.size(); ++s) if (iused[s]) touch_col(it.int_cols[g.c.int_cols[s]], 4);
    if ((which == 4 || which == 7) && !g.is_obj && pr.uses_w) bytes += 8 * (g.k1 - g.k0); // multipliers y
    if (which == 6 && pr.uses_w) bytes += 8 * (g.k1 - g.k0);                                // v[row]
    if (which == 8 && !g.is_obj && pr.uses_w)
      for (int64_t k = g.k0; k < g.k1; ++k) if (!yrow[g.o0 + k]) { yrow[g.o0 + k] = true; bytes += 8; }
  }
  switch (which) {
    case 0: bytes += 8; break;
    case 1: bytes += 8 * P.nvar; break;
    case 2: bytes += 8 * P.loc_ncon; break;
    case 3: bytes += 8 * P.loc_nnzj; break;
    case 4: bytes += 8 * P.loc_nnzh; break;
    case 5: bytes += 8 * P.loc_ncon; break;
    case 6: case 7: bytes += 8 * P.nvar; break; // every entry of the dense result is written once
    case 8: bytes += 8 * (P.loc_ncon + P.loc_nnzj + P.loc_nnzh); break;
  }
  p->bytes_cache[which] = bytes;
  return bytes;
}

int64_t iexa_debug_codegen_source(const iexa_plan *p, char *buf, int64_t cap) {
  if (!p || !p->plan.finalized) return -1;
  const int set = cap < 0 ? 1 : 0; // cap < 0: the product module (jprod! / jtprod! / hprod!), capacity -cap
  if (cap < 0) cap = -cap;
  std::string src = iexa::Specialiser::generate_source(p->plan, set);
  if (buf && cap > 0) {
    size_t n = std::min<size_t>((size_t)cap - 1, src.size());
    std::memcpy(buf, src.data(), n);
    buf[n] = 0;
  }
  return (int64_t)src.size();
}
int64_t iexa_debug_codegen_source_of(const iexa_plan *p, int32_t set, char *buf, int64_t cap) {
  if (!p || !p->plan.finalized || set < 0 || set > 2) return -1;
  std::string src = iexa::Specialiser::generate_source(p->plan, set);
  if (buf && cap > 0) {
    size_t n = std::min<size_t>((size_t)cap - 1, src.size());
    std::memcpy(buf, src.data(), n);
    buf[n] = 0;
  }
  return (int64_t)src.size();
}
int32_t iexa_debug_codegen_compile_of(const iexa_plan *p, int32_t set, int64_t *cubin_bytes) {
  GUARD_BEGIN
  NEED_PLAN(p);
  if (!p->plan.finalized || set < 0 || set > 2) return fail(IEXA_ERR_STATE, "plan not finalized / bad kernel set");
  std::string src = iexa::Specialiser::generate_source(p->plan, set), err;
  std::vector<char> cubin;
  if (!iexa::compile_cubin(src, cubin, err)) return fail(IEXA_ERR_NVRTC, err);
  if (cubin_bytes) *cubin_bytes = (int64_t)cubin.size();
  return IEXA_OK;
  GUARD_END
}
int32_t iexa_debug_set_class_mode(iexa_plan *p, int32_t on) {
  GUARD_BEGIN
  NEED_PLAN(p);
  if (!p->plan.finalized) return fail(IEXA_ERR_STATE, "plan not finalized");
  if (p->engine) return fail(IEXA_ERR_STATE, "only for plans finalized with IEXA_F_NO_DEVICE");
  p->plan.build_groups(on != 0);
  return IEXA_OK;
  GUARD_END
}
int32_t iexa_debug_codegen_compile(const iexa_plan *p, int64_t *cubin_bytes) {
  GUARD_BEGIN
  NEED_PLAN(p);
  if (!p->plan.finalized) return fail(IEXA_ERR_STATE, "plan not finalized");
  const int set = (cubin_bytes && *cubin_bytes == -1) ? 1 : 0; // in: -1 selects the product module
  std::string src = iexa::Specialiser::generate_source(p->plan, set), err;
  std::vector<char> cubin;
  if (!iexa::compile_cubin(src, cubin, err)) return fail(IEXA_ERR_NVRTC, err);
  if (cubin_bytes) *cubin_bytes = (int64_t)cubin.size();
  return IEXA_OK;
  GUARD_END
}

int32_t iexa_debug_cache_stats(int32_t *nvrtc_compiles, int32_t *disk_hits) {
  int a = 0, b = 0;
  iexa::cache_stats(&a, &b);
  if (nvrtc_compiles) *nvrtc_compiles = a;
  if (disk_hits) *disk_hits = b;
  return IEXA_OK;
}

const char *iexa_engine_note(const iexa_plan *p) { return (p && p->engine) ? p->engine->note() : ""; }

int32_t iexa_launches_per_call(const iexa_plan *p, int32_t which) {
  if (!p || !p->engine) return 0;
  return p->engine->launches(which);
}

} // extern "C"

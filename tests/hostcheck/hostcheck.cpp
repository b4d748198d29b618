// hostcheck.cpp — TEST-ONLY host executor of the engine's register programs.
//
// Built (g++, no CUDA runtime) into tests/hostcheck/libiexa_hostcheck.so together with api.cpp
// and codegen.cpp so that `pytest -m "not gpu"` can validate the plan compiler (symbolic
// sparsity, AD programs, layout, byte accounting, NVRTC source generation) on a machine without
// a GPU.  It is NOT part of libiexa_b200.so and is not a fallback: in this library
// make_cuda_engine() always fails, exactly like the product on a GPU-less machine.
#include <cmath>
#include <cstring>

#include "../../include/iexa.h"
#include "../../infiniteexamodels.jl_b200/csrc/api_internal.hpp"
#include "../../infiniteexamodels.jl_b200/csrc/exec.hpp"

namespace iexa {
Engine *make_cuda_engine(Plan &, int, uint32_t, std::string &err) {
  err = "hostcheck build: no CUDA engine";
  return nullptr;
}

// Residency emulation (world > 1): what the CUDA engine keeps on a rank's device — the slices of Plan::column_read_ranges and the
// theta ranges of Plan::theta_read_ranges.  hostcheck_eval_local counts every access outside them (hostcheck_residency_misses).
struct Residency {
  std::vector<std::pair<int64_t, int64_t>> cols, theta;
  int64_t misses = 0;
  void check_col(int32_t col, int64_t j) { if (j < cols[col].first || j >= cols[col].second) ++misses; }
  void check_theta(int64_t i0) {
    for (auto &r : theta) if (i0 >= r.first && i0 < r.second) return;
    ++misses;
  }
};
static Residency *g_res = nullptr;
static int64_t g_last_misses = -1;

static void run_program(const Plan &P, const Generator &g, const Program &pr, int64_t k, const double *x,
                        double W, std::vector<double> &r, double *out, const double *v = nullptr) {
  r.resize(pr.nreg > 0 ? pr.nreg : 1);
  if (g_res) { // every column position and theta entry this support touches must be resident on this rank
    const Iterator &it = P.itrs[g.itr];
    for (int32_t s : g.c.int_cols) { const ColRef &c = it.int_cols[s]; if (!P.columns[c.col].affine) g_res->check_col(c.col, (k / c.div) % c.mod); }
    for (const Instr &I : pr.code) {
      if (I.op == D_FIELD) { const ColRef &c = it.fp_cols[g.c.fp_cols[I.a]]; g_res->check_col(c.col, (k / c.div) % c.mod); }
      if (I.op == D_LOADP) g_res->check_theta(P.index_value(g, I.a, k) - 1);
    }
  }
  for (const Instr &I : pr.code) {
    switch (I.op) {
      case D_FIELD: r[I.dst] = P.fp_col_value(g, I.a, k); break;
      case D_LOADX: r[I.dst] = x[P.index_value(g, I.a, k) - 1]; break;
      case D_LOADP: r[I.dst] = P.theta[P.index_value(g, I.a, k) - 1]; break;
      case D_LOADV: r[I.dst] = v[P.index_value(g, I.a, k) - 1]; break;
      case D_W: r[I.dst] = W; break;
      case D_SEL2: r[I.dst] = P.index_value(g, I.a, k) == P.index_value(g, I.b, k) ? 2.0 : 1.0; break;
      case D_SELNE: r[I.dst] = P.index_value(g, I.a, k) != P.index_value(g, I.b, k) ? 1.0 : 0.0; break;
      case D_OUT: out[I.dst] = I.a >= 0 ? r[I.a] : pr.cpool[~I.a]; break;
      default: {
        double a = I.a >= 0 ? r[I.a] : pr.cpool[~I.a];
        double b = I.b >= 0 ? r[I.b] : pr.cpool[~I.b];
        r[I.dst] = eval_arith(I.op, a, b);
      }
    }
  }
}

// fused group programs (plan.hpp: Group) — slot numbering comes from the group's SlotCtx
static int64_t g_int_col(const Plan &P, const Group &G, int32_t slot, int64_t k) {
  const ColRef &r = P.itrs[G.itr].int_cols[G.ctx.int_cols[slot]];
  int64_t j = (k / r.div) % r.mod;
  const HostColumn &c = P.columns[r.col];
  return c.ival(j);
}
// inst: instance number of a shape-class group (-1: ordinary group)
static int64_t g_index(const Plan &P, const Group &G, int32_t islot, int64_t k, int64_t inst = -1) {
  const IndexExpr &e = G.ctx.uidx[islot];
  int64_t v = inst >= 0 ? G.inst_base[inst * G.ctx.uidx.size() + islot] : e.base; // shape class: the instance's own base
  for (auto &t : e.terms) v += t.second * g_int_col(P, G, t.first, k);
  return v;
}
static void run_group(const Plan &P, const Group &G, const Program &pr, int64_t k, const double *x, const double *y,
                      double sigma, std::vector<double> &r, double *out, int64_t inst = -1, const double *v = nullptr) {
  const std::vector<Generator> &gens_ = G.is_obj ? P.objs : P.cons;
  r.resize(pr.nreg > 0 ? pr.nreg : 1);
  for (const Instr &I : pr.code) {
    switch (I.op) {
      case D_FIELD: { const ColRef &c = P.itrs[G.itr].fp_cols[G.ctx.fp_cols[I.a]]; r[I.dst] = P.columns[c.col].fvals[(k / c.div) % c.mod]; break; }
      case D_LOADX: r[I.dst] = x[g_index(P, G, I.a, k, inst) - 1]; break;
      case D_LOADP: r[I.dst] = P.theta[g_index(P, G, I.a, k, inst) - 1]; break;
      case D_LOADV: r[I.dst] = v[g_index(P, G, I.a, k, inst) - 1]; break;
      case D_SELNE: r[I.dst] = g_index(P, G, I.a, k, inst) != g_index(P, G, I.b, k, inst) ? 1.0 : 0.0; break;
      case D_W: r[I.dst] = G.is_obj ? sigma : (y ? y[(inst >= 0 ? gens_[G.inst_gen(inst, I.a)] : P.member(G, I.a)).o0 + k] : 0.0); break;
      case D_SEL2: r[I.dst] = g_index(P, G, I.a, k, inst) == g_index(P, G, I.b, k, inst) ? 2.0 : 1.0; break;
      case D_CPAR: r[I.dst] = gens_[G.inst_gen(inst, G.cpar_member[I.a])].c.tape[G.cpar_nodes[I.a]].c; break;
      case D_OUT: out[I.dst] = I.a >= 0 ? r[I.a] : pr.cpool[~I.a]; break;
      default: {
        double a = I.a >= 0 ? r[I.a] : pr.cpool[~I.a];
        double b = I.b >= 0 ? r[I.b] : pr.cpool[~I.b];
        r[I.dst] = eval_arith(I.op, a, b);
      }
    }
  }
}
} // namespace iexa

using namespace iexa;

extern "C" {

// which: 0 obj (out[0]), 1 grad (dense nvar), 2 cons, 3 jac_coord, 4 hess_coord  — GLOBAL layout
int32_t hostcheck_eval(iexa_plan *p, int32_t which, const double *x, const double *y, double sigma, double *out) {
  if (!p || !p->plan.finalized) return IEXA_ERR_STATE;
  const Plan &P = p->plan;
  std::vector<double> r, tmp;
  if (which == 0) out[0] = 0.0;
  if (which == 1) std::memset(out, 0, sizeof(double) * (size_t)P.nvar);
  auto each = [&](const Generator &g) {
    const Program &pr = (which == 0 || which == 2) ? g.c.val : (which == 1 || which == 3) ? g.c.d1 : g.c.d2;
    tmp.assign(pr.nout > 0 ? pr.nout : 1, 0.0);
    for (int64_t k = 0; k < g.K; ++k) {
      double W = g.is_obj ? sigma : (y ? y[g.o0 + k] : 0.0);
      run_program(P, g, pr, k, x, W, r, tmp.data());
      switch (which) {
        case 0: out[0] += tmp[0]; break;
        case 1: for (int c = 0; c < g.c.o1step; ++c) out[P.index_value(g, g.c.jac_slot[c], k) - 1] += tmp[c]; break;
        case 2: out[g.o0 + k] = tmp[0]; break;
        case 3: for (int c = 0; c < g.c.o1step; ++c) out[g.o1 + k * g.c.o1step + c] = tmp[c]; break;
        case 4: for (int c = 0; c < g.c.o2step; ++c) out[g.o2 + k * g.c.o2step + c] = tmp[c]; break;
      }
    }
  };
  if (which == 0 || which == 1 || which == 4) for (auto &g : P.objs) each(g);
  if (which >= 2) for (auto &g : P.cons) each(g);
  return IEXA_OK;
}

// this rank's shard only, LOCAL layout (what the CUDA engine produces when world > 1); y is local too.
// which: 0 obj partial, 1 grad partial (dense nvar), 2 cons, 3 jac_coord, 4 hess_coord
int64_t hostcheck_residency_misses(void) { return g_last_misses; }
// out4: [column bytes resident on this rank, column bytes of the whole model, theta bytes read by this rank, 8 * npar]
int32_t hostcheck_residency_bytes(iexa_plan *p, int64_t *out4) {
  if (!p || !p->plan.finalized) return IEXA_ERR_STATE;
  const Plan &P = p->plan;
  auto rr = P.column_read_ranges();
  out4[0] = out4[1] = out4[2] = 0;
  std::vector<char> used(P.columns.size(), 0);
  auto mark = [&](const Generator &g) {
    const Iterator &it = P.itrs[g.itr];
    for (int32_t s : g.c.int_cols) used[it.int_cols[s].col] = 1;
    for (int32_t s : g.c.fp_cols) used[it.fp_cols[s].col] = 1;
  };
  for (auto &g : P.objs) mark(g);
  for (auto &g : P.cons) mark(g);
  for (auto &g : P.pfuncs) mark(g);
  for (size_t c = 0; c < P.columns.size(); ++c) {
    const HostColumn &hc = P.columns[c];
    if (hc.gen_src >= 0 && used[c]) used[hc.gen_src] = 1;
  }
  for (size_t c = 0; c < P.columns.size(); ++c) {
    const HostColumn &hc = P.columns[c];
    if (!used[c] || (hc.is_int && hc.affine)) continue;
    const int64_t w = hc.is_int ? 4 : 8;
    out4[0] += w * std::max<int64_t>(rr[c].second - rr[c].first, 0);
    out4[1] += w * hc.K;
  }
  for (auto &r : P.theta_read_ranges()) out4[2] += 8 * (r.second - r.first);
  out4[3] = 8 * P.npar;
  return IEXA_OK;
}
int32_t hostcheck_eval_local(iexa_plan *p, int32_t which, const double *x, const double *y, double sigma, double *out) {
  if (!p || !p->plan.finalized) return IEXA_ERR_STATE;
  const Plan &P = p->plan;
  Residency res;
  res.cols = P.column_read_ranges();
  res.theta = P.theta_read_ranges();
  g_res = &res;
  struct Done { Residency &r; ~Done() { g_last_misses = r.misses; g_res = nullptr; } } done{res};
  std::vector<double> r, tmp;
  if (which == 0) out[0] = 0.0;
  if (which == 1) std::memset(out, 0, sizeof(double) * (size_t)P.nvar);
  auto each = [&](const Generator &g) {
    const Program &pr = (which == 0 || which == 2) ? g.c.val : (which == 1 || which == 3) ? g.c.d1 : g.c.d2;
    tmp.assign(pr.nout > 0 ? pr.nout : 1, 0.0);
    for (int64_t k = g.k0; k < g.k1; ++k) {
      const int64_t kl = k - g.k0;
      double W = g.is_obj ? sigma : (y ? y[g.l0 + kl] : 0.0);
      run_program(P, g, pr, k, x, W, r, tmp.data());
      switch (which) {
        case 0: out[0] += tmp[0]; break;
        case 1: for (int c = 0; c < g.c.o1step; ++c) out[P.index_value(g, g.c.jac_slot[c], k) - 1] += tmp[c]; break;
        case 2: out[g.l0 + kl] = tmp[0]; break;
        case 3: for (int c = 0; c < g.c.o1step; ++c) out[g.l1 + kl * g.c.o1step + c] = tmp[c]; break;
        case 4: for (int c = 0; c < g.c.o2step; ++c) out[g.l2 + kl * g.c.o2step + c] = tmp[c]; break;
      }
    }
  };
  if (which == 0 || which == 1 || which == 4) for (auto &g : P.objs) each(g);
  if (which >= 2) for (auto &g : P.cons) each(g);
  return IEXA_OK;
}

// same as hostcheck_eval but through the FUSED group programs; returns the number of groups via ngroups
int32_t hostcheck_set_class_mode(iexa_plan *p, int32_t on) {
  if (!p || !p->plan.finalized) return IEXA_ERR_STATE;
  try { p->plan.build_groups(on != 0); } catch (const std::exception &e) { return IEXA_ERR_INVALID; }
  return IEXA_OK;
}

int32_t hostcheck_eval_groups(iexa_plan *p, int32_t which, const double *x, const double *y, double sigma, double *out,
                              int32_t *ngroups) {
  if (!p || !p->plan.finalized) return IEXA_ERR_STATE;
  const Plan &P = p->plan;
  std::vector<double> r, tmp;
  if (which == 0) out[0] = 0.0;
  if (which == 1) std::memset(out, 0, sizeof(double) * (size_t)P.nvar);
  const int prog = (which == 0 || which == 2) ? 0 : (which == 1 || which == 3) ? 1 : 2;
  if (ngroups) *ngroups = (int32_t)P.groups.size();
  for (const Group &G : P.groups) {
    bool want = G.is_obj ? (which == 0 || which == 1 || which == 4) : (which >= 2);
    if (!want) continue;
    const Program &pr = G.prog[prog];
    tmp.assign(pr.nout > 0 ? pr.nout : 1, 0.0);
    const std::vector<Generator> &gens = G.is_obj ? P.objs : P.cons;
    const size_t ninst = G.is_class ? G.n_inst() : 1;
    for (size_t ii = 0; ii < ninst; ++ii)
    for (int64_t k = 0; k < G.K; ++k) {
      const int64_t inst = G.is_class ? (int64_t)ii : -1;
      run_group(P, G, pr, k, x, y, sigma, r, tmp.data(), inst);
      for (size_t j = 0; j < G.outmap[prog].size(); ++j) {
        int m = G.outmap[prog][j].first, c = G.outmap[prog][j].second;
        const Generator &g = inst >= 0 ? gens[G.inst_gen(ii, m)] : P.member(G, m);
        switch (which) {
          case 0: out[0] += tmp[j]; break;
          case 1: out[g_index(P, G, G.jac_slot[m][c], k, inst) - 1] += tmp[j]; break;
          case 2: out[g.o0 + k] = tmp[j]; break;
          case 3: out[g.o1 + k * g.c.o1step + c] = tmp[j]; break;
          case 4: out[g.o2 + k * g.c.o2step + c] = tmp[j]; break;
        }
      }
    }
  }
  return IEXA_OK;
}

// matrix-free products through the per-generator programs (use_groups == 0) or the fused group programs.
// which: 5 jprod (out: ncon), 6 jtprod (v: ncon rows, out: nvar), 7 hprod (out: nvar).  GLOBAL layout.
// phase_stats (optional, 4 int64): groups in phase 0 / direct outputs / zero-filled entries / groups with work
int32_t hostcheck_prod(iexa_plan *p, int32_t which, int32_t use_groups, const double *x, const double *y, const double *v,
                       double sigma, double *out, int64_t *phase_stats) {
  if (!p || !p->plan.finalized || which < 5 || which > 7) return IEXA_ERR_STATE;
  const Plan &P = p->plan;
  std::vector<double> r, tmp;
  const int64_t no = which == 5 ? P.ncon : P.nvar;
  const double NaN = std::nan("");
  // emulate the device protocol: out starts as garbage; zero ranges are zeroed; phase-0 direct outputs STORE
  for (int64_t i = 0; i < no; ++i) out[i] = use_groups && which != 5 ? NaN : 0.0;
  if (!use_groups) {
    auto each = [&](const Generator &g) {
      const Program &pr = which == 5 ? g.c.jv : which == 6 ? g.c.jtv : g.c.hv;
      if (g.is_obj && which != 7) return;
      tmp.assign(pr.nout > 0 ? pr.nout : 1, 0.0);
      for (int64_t k = 0; k < g.K; ++k) {
        // J'v: the root weight is v[row]
        double W = which == 6 ? v[g.o0 + k] : (g.is_obj ? sigma : (y ? y[g.o0 + k] : 0.0));
        run_program(P, g, pr, k, x, W, r, tmp.data(), v);
        if (which == 5) out[g.o0 + k] = tmp[0];
        else {
          const std::vector<int32_t> &sl = which == 6 ? g.c.jtv_slot : g.c.hv_slot;
          for (int j = 0; j < pr.nout; ++j) out[P.index_value(g, sl[j], k) - 1] += tmp[j];
        }
      }
    };
    for (auto &g : P.objs) each(g);
    for (auto &g : P.cons) each(g);
    return IEXA_OK;
  }
  const int prog = which - 2, w = which - 6;
  int64_t st[4] = {0, 0, 0, 0};
  if (which != 5) for (auto &z : P.scat_zero_ranges[w]) { for (int64_t i = 0; i < z.second; ++i) out[z.first + i] = 0.0; st[2] += z.second; }
  // device protocol: phase-0 primaries, then their riders (same thread, right after the primary's body), then phase 1
  for (int phase = 0; phase < 3; ++phase)
  for (const Group &G : P.groups) {
    if (G.is_obj && which != 7) continue;
    const Program &pr = G.prog[prog];
    if (which != 5) {
      if (pr.nout == 0) continue;
      const bool is_rider = G.scat_phase[w] == 0 && G.scat_rider_of[w] >= 0;
      const int gp = G.scat_phase[w] == 1 ? 2 : (is_rider ? 1 : 0);
      if (gp != phase) continue;
    }
    if (which == 5 && phase != 0) continue;
    if (which != 5) { st[3]++; if (phase == 0) st[0]++; }
    tmp.assign(pr.nout > 0 ? pr.nout : 1, 0.0);
    const std::vector<Generator> &gens = G.is_obj ? P.objs : P.cons;
    const size_t ninst = G.is_class ? G.n_inst() : 1;
    // J'v: D_W(member) reads v[row of member] — hand v in as y
    const double *wy = which == 6 ? v : y;
    for (size_t ii = 0; ii < ninst; ++ii)
    for (int64_t k = 0; k < G.K; ++k) {
      const int64_t inst = G.is_class ? (int64_t)ii : -1;
      run_group(P, G, pr, k, x, wy, sigma, r, tmp.data(), inst, v);
      for (size_t j = 0; j < G.outmap[prog].size(); ++j) {
        if (which == 5) { const Generator &g = inst >= 0 ? gens[G.inst_gen(ii, G.outmap[prog][j].first)] : P.member(G, G.outmap[prog][j].first); out[g.o0 + k] = tmp[j]; continue; }
        const int64_t i = g_index(P, G, G.outmap[prog][j].second, k, inst) - 1;
        const int mode = phase < 2 ? G.scat_direct[w][j] : 0;
        if (mode == 1) { out[i] = tmp[j]; if (ii == 0 && k == 0) st[1]++; } else out[i] += tmp[j]; // mode 2 (rider): += onto the primary's store
      }
    }
  }
  if (phase_stats) for (int i = 0; i < 4; ++i) phase_stats[i] = st[i];
  return IEXA_OK;
}

// two-phase scatter analysis of jtprod! (w = 0) / hprod! (w = 1): out = {phase-0 primaries, riders, phase-1 groups}
int32_t hostcheck_scatter_info(iexa_plan *p, int32_t w, int64_t *out) {
  if (!p || !p->plan.finalized || w < 0 || w > 1) return IEXA_ERR_STATE;
  out[0] = out[1] = out[2] = 0;
  for (const Group &G : p->plan.groups) {
    if (G.prog[4 + w].nout == 0) continue;
    if (G.scat_phase[w] == 1) out[2]++;
    else if (G.scat_rider_of[w] >= 0) out[1]++;
    else out[0]++;
  }
  return IEXA_OK;
}

// which: 0 jac, 1 hess; 1-based int64 rows/cols, GLOBAL layout
int32_t hostcheck_structure(iexa_plan *p, int32_t which, int64_t *rows, int64_t *cols) {
  if (!p || !p->plan.finalized) return IEXA_ERR_STATE;
  const Plan &P = p->plan;
  auto each = [&](const Generator &g) {
    for (int64_t k = 0; k < g.K; ++k) {
      if (which == 0) {
        for (int c = 0; c < g.c.o1step; ++c) {
          rows[g.o1 + k * g.c.o1step + c] = g.o0 + k + 1;
          cols[g.o1 + k * g.c.o1step + c] = P.index_value(g, g.c.jac_slot[c], k);
        }
      } else {
        for (int c = 0; c < g.c.o2step; ++c) {
          int64_t i = P.index_value(g, g.c.hess_slot[c].first, k), j = P.index_value(g, g.c.hess_slot[c].second, k);
          rows[g.o2 + k * g.c.o2step + c] = i >= j ? i : j;
          cols[g.o2 + k * g.c.o2step + c] = i >= j ? j : i;
        }
      }
    }
  };
  if (which == 1) for (auto &g : P.objs) each(g);
  for (auto &g : P.cons) each(g);
  return IEXA_OK;
}

// grad! analysis: out[0] = number of single-writer slots, out[1] = entries grad! must zero-fill
int32_t hostcheck_grad_stats(iexa_plan *p, int64_t *out) {
  if (!p || !p->plan.finalized) return IEXA_ERR_STATE;
  out[0] = out[1] = 0;
  for (auto &g : p->plan.objs) for (uint8_t d : g.grad_direct) out[0] += d;
  for (auto &z : p->plan.grad_zero_ranges) out[1] += z.second;
  return IEXA_OK;
}

// per-generator compile statistics: out[8] = {o1step, o2step, n_occ1, n_occ2, nreg_val, nreg_d1, nreg_d2, ncode_d2}
int32_t hostcheck_gen_stats(iexa_plan *p, int32_t is_obj, int32_t i, int64_t *out) {
  if (!p) return IEXA_ERR_INVALID;
  const auto &v = is_obj ? p->plan.objs : p->plan.cons;
  if (i < 0 || i >= (int)v.size()) return IEXA_ERR_INVALID;
  const Generator &g = v[i];
  out[0] = g.c.o1step; out[1] = g.c.o2step; out[2] = g.c.n_occ1; out[3] = g.c.n_occ2;
  out[4] = g.c.val.nreg; out[5] = g.c.d1.nreg; out[6] = g.c.d2.nreg; out[7] = (int64_t)g.c.d2.code.size();
  return IEXA_OK;
}
}

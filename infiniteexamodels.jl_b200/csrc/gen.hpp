// gen.hpp — compile ONE generator (expression tree × iterator) into sparsity metadata and
// register programs.
//
// What is restated here (ExaModels 0.11.2, not vendored in /root/reference; see SURVEY.md
// Appendix A — the slot-ORDER policy below is a documented hypothesis until dumps from a
// real ExaModels exist, "parity unpinned"):
//   * first-order occurrences are the Var leaves in left-to-right order; leaves whose index
//     EXPRESSIONS are identical share a slot (o1step = #distinct), first occurrence wins;
//   * second-order occurrences come from the hrpass0 / hrpass / hdrpass recursion:
//     top-level +,-,const* chains emit nothing; below the first nonlinear node every leaf
//     emits a diagonal slot and every binary node a cross walk over its two subtrees;
//   * slots are laid out per support: slot = o + ostep*(k-1) + c.
// The tape comes from the lowering of src/transform.jl:337-389 (_exafy) and :290-334
// (_map_variable).
#pragma once
#include <algorithm>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>

#include "../../include/iexa.h"
#include "dag.hpp"

namespace iexa {

struct IndexExpr {
  int64_t base = 0;
  std::vector<std::pair<int32_t, int64_t>> terms; // (int column, coef), sorted, coef != 0
  bool operator==(const IndexExpr &o) const { return base == o.base && terms == o.terms; }
};

// slot numbering shared by everything compiled into one Dag: a single generator, or a GROUP of
// generators over the same iterator that are fused into one program (plan.hpp: Group)
struct SlotCtx {
  std::vector<IndexExpr> uidx;   // distinct canonical index expressions ("index slots")
  std::vector<int32_t> fp_cols;  // fp column slot -> iterator fp column
  std::vector<int32_t> int_cols; // int column slot -> iterator int column (terms are rewritten to slots)
  std::map<int32_t, int32_t> icolslot, fcolslot;
};

struct GenCompiled {
  // inputs
  std::vector<iexa_node> tape;
  std::vector<iexa_index> raw_idx;
  std::vector<IndexExpr> uidx;   // copy of the SlotCtx the generator was compiled against
  std::vector<int32_t> idx_map;  // caller's index id -> index slot
  std::vector<int32_t> fp_cols;
  std::vector<int32_t> int_cols;
  // sparsity (symbolic)
  std::vector<int32_t> jac_slot;                      // per first-order slot: index slot of the variable
  std::vector<std::pair<int32_t, int32_t>> hess_slot; // per second-order slot: (index slot, index slot)
  int32_t o1step = 0, o2step = 0;
  int32_t n_occ1 = 0, n_occ2 = 0; // occurrences before compression (reporting)
  bool is_null = false;           // constant-only generator (ExaModels.Null, transform.jl:393)
  std::vector<int32_t> jac_perm;  // IEXA_SLOT_ORDER_JAC_ROW_SORTED: applied permutation of the first-order slots (empty: policy order kept)
  bool jac_row_sorted = false;    // the slots of every row are in strictly increasing column order
  // programs
  Program val, d1, d2;
  std::vector<uint8_t> x_slots_val, x_slots_d1, x_slots_d2; // which index slots each program LOADX-es
  // matrix-free products (jprod! / jtprod! / hprod!) as programs of their own: jv has ONE output (the row's
  // sum_c d1[c]*v[col_c]); jtv / hv have one output per touched index slot (jtv_slot / hv_slot name it)
  Program jv, jtv, hv;
  std::vector<int32_t> jtv_slot, hv_slot;
  std::vector<uint8_t> x_slots_prod; // scratch of schedule()
};

// ---- matrix-free products as DAG outputs --------------------------------------------------------------
// ExaModels reuses its reverse passes with an (out, v) pair in place of the value vector (SURVEY App. A.5:
// Jv[row] += adj*v[col], Jtv[col] += adj*v[row], Hv[i] += h*v[j] (+ mirror)).  Here the products are built ONCE from
// the symbolic first / second order slot values of the members of a Dag (a generator, or a fused group): no COO
// values are materialised, and everything the slots share with each other (trig of the same state, common adjoints)
// is shared by the product too.
struct MemberSlots {
  std::vector<int> s1, s2;                              // DAG ids of the first / second order slot values
  std::vector<int32_t> jac_slot;                        // index slot per first-order slot
  std::vector<std::pair<int32_t, int32_t>> hess_slot;   // index-slot pair per second-order slot
  int wid = 0;                                          // member id of the root weight W (D_W)
  int val = -1;                                         // DAG id of the member's value
};
struct ProductOuts {
  std::vector<int> jv;                    // one node per member
  std::vector<int> jtv, hv;               // one node per touched index slot, in order of first appearance
  std::vector<int32_t> jtv_slot, hv_slot;
};
inline ProductOuts build_products(Dag &d, const std::vector<MemberSlots> &M, bool is_obj) {
  ProductOuts P;
  std::map<int32_t, int> jpos, hpos;
  auto acc = [&](std::map<int32_t, int> &pos, std::vector<int> &outs, std::vector<int32_t> &slots, int32_t u, int term) {
    auto it = pos.find(u);
    if (it == pos.end()) { it = pos.emplace(u, (int)outs.size()).first; outs.push_back(d.cnst(0.0)); slots.push_back(u); }
    outs[it->second] = d.add(outs[it->second], term);
  };
  for (const MemberSlots &m : M) {
    if (!is_obj) {
      int row = d.cnst(0.0);
      for (size_t c = 0; c < m.s1.size(); ++c) row = d.add(row, d.mul(m.s1[c], d.loadv(m.jac_slot[c])));
      P.jv.push_back(row);
      // J'v: the weight leaf W of member wid is v[row] here (the kernel is handed v in place of y)
      for (size_t c = 0; c < m.s1.size(); ++c) acc(jpos, P.jtv, P.jtv_slot, m.jac_slot[c], d.mul(m.s1[c], d.w(m.wid)));
    }
    // Hv: slot (u1, u2) with value h (root weight already inside) is the lower-triangle entry (max, min) of the
    // symmetric matrix: Hv[u1] += h*v[u2], and the mirror Hv[u2] += h*v[u1] unless both land on the same variable
    for (size_t c = 0; c < m.s2.size(); ++c) {
      const int32_t u1 = m.hess_slot[c].first, u2 = m.hess_slot[c].second;
      acc(hpos, P.hv, P.hv_slot, u1, d.mul(m.s2[c], d.loadv(u2)));
      if (u1 != u2) acc(hpos, P.hv, P.hv_slot, u2, d.mul(d.mul(m.s2[c], d.selne(u1, u2)), d.loadv(u1)));
    }
  }
  return P;
}

Program schedule(const Dag &dag, const std::vector<int> &outs, size_t n_islots, std::vector<uint8_t> &x_slots);

class GenCompiler {
 public:
  // standalone: own slot context and Dag
  GenCompiler(const iexa_node *nodes, int32_t n, const iexa_index *idx, int32_t n_idx,
              int32_t n_int_cols_itr, int32_t n_fp_cols_itr)
      : n_(n), n_int_itr_(n_int_cols_itr), n_fp_itr_(n_fp_cols_itr), ctx_(own_ctx_), dag_(own_dag_), wid_(0) {
    init(nodes, idx, n_idx);
  }
  // fused: shared slot context and Dag; wid = member id used for the root weight W
  GenCompiler(const iexa_node *nodes, int32_t n, const iexa_index *idx, int32_t n_idx,
              int32_t n_int_cols_itr, int32_t n_fp_cols_itr, SlotCtx &ctx, Dag &dag, int wid)
      : n_(n), n_int_itr_(n_int_cols_itr), n_fp_itr_(n_fp_cols_itr), ctx_(ctx), dag_(dag), wid_(wid) {
    init(nodes, idx, n_idx);
  }

  GenCompiled g;
  // shape classes: tape node -> per-instance constant parameter index (or -1 = literal); see plan.hpp
  const std::vector<int32_t> *cpar_of_node = nullptr;

  // symbolic AD into the Dag; afterwards val_root()/slot1()/slot2() name the outputs
  void differentiate() {
    build_values();
    first_order();
    if (jac_perm) apply_jac_perm(*jac_perm);
    second_order();
  }
  // IEXA_SLOT_ORDER_JAC_ROW_SORTED: first-order slot c of the policy order becomes slot position p with perm[p] = c (the
  // slots of a row sorted by column index — computed by Plan::jac_row_sort, which knows the iterator's columns)
  const std::vector<int32_t> *jac_perm = nullptr;
  const SlotCtx &ctx() const { return ctx_; }
  void apply_jac_perm(const std::vector<int32_t> &perm) {
    if (perm.size() != g.jac_slot.size()) return;
    std::vector<int> s1(perm.size());
    std::vector<int32_t> js(perm.size());
    for (size_t p = 0; p < perm.size(); ++p) { s1[p] = slot1_[perm[p]]; js[p] = g.jac_slot[perm[p]]; }
    slot1_.swap(s1);
    g.jac_slot.swap(js);
    slot1_of_.clear();
    for (size_t p = 0; p < g.jac_slot.size(); ++p) slot1_of_[g.jac_slot[p]] = (int32_t)p;
  }
  int val_root() const { return val_[n_ - 1]; }
  const std::vector<int> &slot1() const { return slot1_; }
  const std::vector<int> &slot2() const { return slot2_; }

  void compile() {
    differentiate();
    g.uidx = ctx_.uidx; g.fp_cols = ctx_.fp_cols; g.int_cols = ctx_.int_cols;
    g.val = schedule(dag_, {val_[n_ - 1]}, ctx_.uidx.size(), g.x_slots_val);
    g.d1 = schedule(dag_, slot1_, ctx_.uidx.size(), g.x_slots_d1);
    g.d2 = schedule(dag_, slot2_, ctx_.uidx.size(), g.x_slots_d2);
    MemberSlots ms;
    ms.s1 = slot1_; ms.s2 = slot2_; ms.jac_slot = g.jac_slot; ms.hess_slot = g.hess_slot; ms.wid = wid_;
    ProductOuts po = build_products(dag_, {ms}, is_obj_);
    if (!is_obj_) {
      g.jv = schedule(dag_, po.jv, ctx_.uidx.size(), g.x_slots_prod);
      g.jtv = schedule(dag_, po.jtv, ctx_.uidx.size(), g.x_slots_prod);
      g.jtv_slot = po.jtv_slot;
    }
    g.hv = schedule(dag_, po.hv, ctx_.uidx.size(), g.x_slots_prod);
    g.hv_slot = po.hv_slot;
  }
  bool is_obj_ = false; // objective generators have no rows: no jv / jtv programs
  // Slot-ORDER policy (iexa_set_option IEXA_OPT_SLOT_ORDER).  The order in which the reverse passes meet the Var leaves
  // decides which COO slot a variable (pair) gets; ExaModels' order is a hypothesis here until a dump of the real package
  // exists ("parity unpinned") — so it is DATA: 0 = children left to right (inner1 then inner2), 1 = right
  // to left (inner2 then inner1; cross pairs (inner2-leaf, inner1-leaf)).  Oracle and product honour the same value.
  int order_ = 0;
  void set_options(int slot_order, bool strict) {
    order_ = slot_order;
    dag_.strict = strict;
    if (const char *e = getenv("IEXA_STRICT_IEEE")) dag_.strict = e[0] != '0'; // experiments: override the plan's choice
  }

 private:
  void init(const iexa_node *nodes, const iexa_index *idx, int32_t n_idx) {
    if (n_ <= 0) throw std::invalid_argument("empty tape");

    g.tape.assign(nodes, nodes + n_);
    if (n_idx > 0) g.raw_idx.assign(idx, idx + n_idx);
    canon_indices(idx, n_idx);
    analyse();
  }
  SlotCtx own_ctx_;
  Dag own_dag_;
  enum Kind { K_CONSTCLASS = 0, K_VAR = 1, K_UN = 2, K_BIN = 3 };
  enum Pass { P_NONE = 0, P_KEEP, P_FLIP, P_SCALE, P_ADD, P_SUB };
  struct Info {
    int kind = K_CONSTCLASS;
    int c1 = -1, c2 = -1; // active children (tape ids)
    int pass = P_NONE;    // hrpass0 pass-through class
    int y1 = -1, y2 = -1, h11 = -1, h12 = -1, h22 = -1; // DAG ids of local partials
  };
  int32_t n_, n_int_itr_, n_fp_itr_;
  std::vector<Info> info_;
  std::vector<int> val_;
  SlotCtx &ctx_;
  Dag &dag_;
  int wid_;
  std::vector<int> slot1_, slot2_;
  std::map<int32_t, int32_t> slot1_of_;
  std::map<std::pair<int32_t, int32_t>, int32_t> slot2_of_;

  static bool is_leaf(int op) { return op >= IEXA_OP_CONST && op <= IEXA_OP_PAR; }
  static bool is_bin(int op) { return op >= IEXA_OP_ADD && op <= IEXA_OP_POW; }
  static bool is_un(int op) { return op >= IEXA_OP_NEG && op < IEXA_OP__END; }

  void canon_indices(const iexa_index *idx, int32_t n_idx) {
    // which int columns are referenced (stable order of first reference)
    for (int32_t i = 0; i < n_idx; ++i) {
      if (idx[i].nterms < 0 || idx[i].nterms > IEXA_MAX_INDEX_TERMS)
        throw std::invalid_argument("index expression: bad nterms");
      for (int t = 0; t < idx[i].nterms; ++t) {
        int32_t c = idx[i].col[t];
        if (c < 0 || c >= n_int_itr_) throw std::invalid_argument("index expression: int column out of range");
        if (idx[i].coef[t] != 0 && !ctx_.icolslot.count(c)) {
          ctx_.icolslot[c] = (int32_t)ctx_.int_cols.size();
          ctx_.int_cols.push_back(c);
        }
      }
    }
    g.idx_map.resize(n_idx);
    for (int32_t i = 0; i < n_idx; ++i) {
      std::map<int32_t, int64_t> acc;
      for (int t = 0; t < idx[i].nterms; ++t)
        if (idx[i].coef[t] != 0) acc[ctx_.icolslot[idx[i].col[t]]] += idx[i].coef[t];
      IndexExpr e;
      e.base = idx[i].base;
      for (auto &kv : acc)
        if (kv.second != 0) e.terms.push_back(kv);
      int32_t found = -1;
      for (size_t u = 0; u < ctx_.uidx.size(); ++u)
        if (ctx_.uidx[u] == e) { found = (int32_t)u; break; }
      if (found < 0) { found = (int32_t)ctx_.uidx.size(); ctx_.uidx.push_back(e); }
      g.idx_map[i] = found;
    }
  }

  void analyse() {
    info_.assign(n_, Info());
    for (int i = 0; i < n_; ++i) {
      const iexa_node &nd = g.tape[i];
      Info &in = info_[i];
      if (is_leaf(nd.op)) {
        if (nd.op == IEXA_OP_VAR || nd.op == IEXA_OP_PAR) {
          if (nd.a < 0 || nd.a >= (int)g.idx_map.size())
            throw std::invalid_argument("tape: index-expression id out of range");
        }
        if (nd.op == IEXA_OP_FIELD) {
          if (nd.a < 0 || nd.a >= n_fp_itr_) throw std::invalid_argument("tape: fp column out of range");
          if (!ctx_.fcolslot.count(nd.a)) { ctx_.fcolslot[nd.a] = (int32_t)ctx_.fp_cols.size(); ctx_.fp_cols.push_back(nd.a); }
        }
        in.kind = nd.op == IEXA_OP_VAR ? K_VAR : K_CONSTCLASS;
      } else if (is_bin(nd.op)) {
        if (nd.a < 0 || nd.a >= i || nd.b < 0 || nd.b >= i) throw std::invalid_argument("tape: child id not before parent");
        bool a1 = info_[nd.a].kind != K_CONSTCLASS, a2 = info_[nd.b].kind != K_CONSTCLASS;
        if (a1 && a2) {
          in.kind = K_BIN; in.c1 = nd.a; in.c2 = nd.b;
          in.pass = nd.op == IEXA_OP_ADD ? P_ADD : nd.op == IEXA_OP_SUB ? P_SUB : P_NONE;
        } else if (a1 || a2) {
          in.kind = K_UN; in.c1 = a1 ? nd.a : nd.b;
          if (nd.op == IEXA_OP_ADD) in.pass = P_KEEP;
          else if (nd.op == IEXA_OP_SUB) in.pass = a1 ? P_KEEP : P_FLIP;
          else if (nd.op == IEXA_OP_MUL) in.pass = P_SCALE;
        }
      } else if (is_un(nd.op)) {
        if (nd.a < 0 || nd.a >= i) throw std::invalid_argument("tape: child id not before parent");
        if (info_[nd.a].kind != K_CONSTCLASS) {
          in.kind = K_UN; in.c1 = nd.a;
          in.pass = nd.op == IEXA_OP_NEG ? P_FLIP : nd.op == IEXA_OP_POS ? P_KEEP : P_NONE;
        }
      } else {
        throw std::invalid_argument("tape: unsupported operator " + std::to_string(nd.op));
      }
    }
    g.is_null = info_[n_ - 1].kind == K_CONSTCLASS;
  }

  // value + local partials of a unary function u -> f(u); returns value, sets y,h when active
  int lower_unary(int op, int u, bool active, int &y, int &h) {
    Dag &d = dag_;
    const double D2R = 0.017453292519943295769, R2D = 57.295779513082320877;
    int v = -1;
    auto one = [&] { return d.cnst(1.0); };
    switch (op) {
      case IEXA_OP_NEG: v = d.neg(u); if (active) { y = d.cnst(-1.0); h = d.cnst(0.0); } break;
      case IEXA_OP_POS: v = u; if (active) { y = one(); h = d.cnst(0.0); } break;
      case IEXA_OP_INV: v = d.div(one(), u); if (active) { y = d.neg(d.sq(v)); h = d.mul(d.cnst(2.0), d.mul(v, d.sq(v))); } break;
      case IEXA_OP_SQRT: v = d.un(D_SQRT, u); if (active) { y = d.div(d.cnst(0.5), v); h = d.div(d.mul(d.cnst(-0.5), y), u); } break;
      case IEXA_OP_CBRT: v = d.un(D_CBRT, u); if (active) { y = d.div(v, d.mul(d.cnst(3.0), u)); h = d.div(d.mul(d.cnst(-2.0 / 3.0), y), u); } break;
      case IEXA_OP_ABS: v = d.un(D_ABS, u); if (active) { y = d.un(D_SIGNP, u); h = d.cnst(0.0); } break;
      case IEXA_OP_ABS2: v = d.sq(u); if (active) { y = d.mul(d.cnst(2.0), u); h = d.cnst(2.0); } break;
      case IEXA_OP_EXP: v = d.un(D_EXP, u); if (active) { y = v; h = v; } break;
      case IEXA_OP_EXP2: v = d.un(D_EXP2, u); if (active) { y = d.mul(v, d.cnst(M_LN2)); h = d.mul(y, d.cnst(M_LN2)); } break;
      case IEXA_OP_LOG: v = d.un(D_LOG, u); if (active) { y = d.div(one(), u); h = d.neg(d.sq(y)); } break;
      case IEXA_OP_LOG2: v = d.un(D_LOG2, u); if (active) { y = d.div(d.cnst(1.0 / M_LN2), u); h = d.neg(d.div(y, u)); } break;
      case IEXA_OP_LOG10: v = d.un(D_LOG10, u); if (active) { y = d.div(d.cnst(1.0 / M_LN10), u); h = d.neg(d.div(y, u)); } break;
      case IEXA_OP_LOG1P: v = d.un(D_LOG1P, u); if (active) { y = d.div(one(), d.add(one(), u)); h = d.neg(d.sq(y)); } break;
      case IEXA_OP_SIN: v = d.un(D_SIN, u); if (active) { y = d.un(D_COS, u); h = d.neg(v); } break;
      case IEXA_OP_COS: v = d.un(D_COS, u); if (active) { y = d.neg(d.un(D_SIN, u)); h = d.neg(v); } break;
      case IEXA_OP_TAN: v = d.un(D_TAN, u); if (active) { y = d.add(one(), d.sq(v)); h = d.mul(d.mul(d.cnst(2.0), v), y); } break;
      case IEXA_OP_ASIN: v = d.un(D_ASIN, u); if (active) { y = d.div(one(), d.un(D_SQRT, d.sub(one(), d.sq(u)))); h = d.mul(u, d.mul(y, d.sq(y))); } break;
      case IEXA_OP_ACOS: v = d.un(D_ACOS, u); if (active) { y = d.neg(d.div(one(), d.un(D_SQRT, d.sub(one(), d.sq(u))))); h = d.mul(u, d.mul(y, d.sq(y))); } break;
      case IEXA_OP_ATAN: v = d.un(D_ATAN, u); if (active) { y = d.div(one(), d.add(one(), d.sq(u))); h = d.mul(d.mul(d.cnst(-2.0), u), d.sq(y)); } break;
      case IEXA_OP_ACOT: v = d.un(D_ATAN, d.div(one(), u)); if (active) { y = d.neg(d.div(one(), d.add(one(), d.sq(u)))); h = d.mul(d.mul(d.cnst(2.0), u), d.sq(y)); } break;
      case IEXA_OP_CSC: { v = d.div(one(), d.un(D_SIN, u)); if (active) { y = d.neg(d.mul(d.sq(v), d.un(D_COS, u))); h = d.mul(v, d.sub(d.mul(d.cnst(2.0), d.sq(v)), one())); } break; }
      case IEXA_OP_SEC: { v = d.div(one(), d.un(D_COS, u)); if (active) { y = d.mul(d.sq(v), d.un(D_SIN, u)); h = d.mul(v, d.sub(d.mul(d.cnst(2.0), d.sq(v)), one())); } break; }
      case IEXA_OP_COT: { v = d.div(one(), d.un(D_TAN, u)); if (active) { y = d.neg(d.add(one(), d.sq(v))); h = d.mul(d.mul(d.cnst(-2.0), v), y); } break; }
      case IEXA_OP_SINH: v = d.un(D_SINH, u); if (active) { y = d.un(D_COSH, u); h = v; } break;
      case IEXA_OP_COSH: v = d.un(D_COSH, u); if (active) { y = d.un(D_SINH, u); h = v; } break;
      case IEXA_OP_TANH: v = d.un(D_TANH, u); if (active) { y = d.sub(one(), d.sq(v)); h = d.mul(d.mul(d.cnst(-2.0), v), y); } break;
      case IEXA_OP_CSCH: { v = d.div(one(), d.un(D_SINH, u)); if (active) { y = d.neg(d.mul(d.sq(v), d.un(D_COSH, u))); h = d.mul(v, d.add(d.mul(d.cnst(2.0), d.sq(v)), one())); } break; }
      case IEXA_OP_SECH: { v = d.div(one(), d.un(D_COSH, u)); if (active) { y = d.neg(d.mul(d.sq(v), d.un(D_SINH, u))); h = d.mul(v, d.sub(one(), d.mul(d.cnst(2.0), d.sq(v)))); } break; }
      case IEXA_OP_COTH: { v = d.div(one(), d.un(D_TANH, u)); if (active) { y = d.sub(one(), d.sq(v)); h = d.mul(d.mul(d.cnst(-2.0), v), y); } break; }
      case IEXA_OP_ATANH: v = d.un(D_ATANH, u); if (active) { y = d.div(one(), d.sub(one(), d.sq(u))); h = d.mul(d.mul(d.cnst(2.0), u), d.sq(y)); } break;
      case IEXA_OP_ACOTH: v = d.un(D_ATANH, d.div(one(), u)); if (active) { y = d.div(one(), d.sub(one(), d.sq(u))); h = d.mul(d.mul(d.cnst(2.0), u), d.sq(y)); } break;
      // degree variants: g(u * pi/180)  (or 180/pi * g(u) for the inverse functions)
      case IEXA_OP_SIND: case IEXA_OP_COSD: case IEXA_OP_TAND: case IEXA_OP_CSCD: case IEXA_OP_SECD: case IEXA_OP_COTD: {
        static const int base[6] = {IEXA_OP_SIN, IEXA_OP_COS, IEXA_OP_TAN, IEXA_OP_CSC, IEXA_OP_SEC, IEXA_OP_COT};
        int which = op == IEXA_OP_SIND ? 0 : op == IEXA_OP_COSD ? 1 : op == IEXA_OP_TAND ? 2 : op == IEXA_OP_CSCD ? 3 : op == IEXA_OP_SECD ? 4 : 5;
        int ur = d.mul(u, d.cnst(D2R));
        int yy = -1, hh = -1;
        v = lower_unary(base[which], ur, active, yy, hh);
        if (active) { y = d.mul(yy, d.cnst(D2R)); h = d.mul(hh, d.cnst(D2R * D2R)); }
        break;
      }
      case IEXA_OP_ATAND: case IEXA_OP_ACOTD: {
        int yy = -1, hh = -1;
        int vv = lower_unary(op == IEXA_OP_ATAND ? IEXA_OP_ATAN : IEXA_OP_ACOT, u, active, yy, hh);
        v = d.mul(vv, d.cnst(R2D));
        if (active) { y = d.mul(yy, d.cnst(R2D)); h = d.mul(hh, d.cnst(R2D)); }
        break;
      }
      default: throw std::invalid_argument("tape: unsupported unary operator " + std::to_string(op));
    }
    return v;
  }

  void build_values() {
    Dag &d = dag_;
    val_.assign(n_, -1);
    for (int i = 0; i < n_; ++i) {
      const iexa_node &nd = g.tape[i];
      Info &in = info_[i];
      switch (nd.op) {
        case IEXA_OP_CONST:
          val_[i] = (cpar_of_node && (*cpar_of_node)[i] >= 0) ? d.cpar((*cpar_of_node)[i]) : d.cnst(nd.c);
          continue;
        case IEXA_OP_FIELD: val_[i] = d.field(ctx_.fcolslot[nd.a]); continue;
        case IEXA_OP_VAR: val_[i] = d.loadx(g.idx_map[nd.a]); continue;
        case IEXA_OP_PAR: val_[i] = d.loadp(g.idx_map[nd.a]); continue;
        default: break;
      }
      if (is_un(nd.op)) {
        val_[i] = lower_unary(nd.op, val_[nd.a], in.kind == K_UN, in.y1, in.h11);
        continue;
      }
      int u1 = val_[nd.a], u2 = val_[nd.b];
      int v = -1;
      switch (nd.op) {
        case IEXA_OP_ADD: v = d.add(u1, u2); break;
        case IEXA_OP_SUB: v = d.sub(u1, u2); break;
        case IEXA_OP_MUL: v = d.mul(u1, u2); break;
        case IEXA_OP_DIV: v = d.div(u1, u2); break;
        case IEXA_OP_POW: v = d.pow(u1, u2); break;
      }
      val_[i] = v;
      int one = d.cnst(1.0), zero = d.cnst(0.0);
      if (in.kind == K_BIN) {
        switch (nd.op) {
          case IEXA_OP_ADD: in.y1 = one; in.y2 = one; in.h11 = in.h12 = in.h22 = zero; break;
          case IEXA_OP_SUB: in.y1 = one; in.y2 = d.cnst(-1.0); in.h11 = in.h12 = in.h22 = zero; break;
          case IEXA_OP_MUL: in.y1 = u2; in.y2 = u1; in.h11 = zero; in.h12 = one; in.h22 = zero; break;
          case IEXA_OP_DIV:
            in.y1 = d.div(one, u2);
            in.y2 = d.neg(d.div(v, u2));
            in.h11 = zero;
            in.h12 = d.neg(d.sq(in.y1));
            in.h22 = d.div(d.mul(d.cnst(-2.0), in.y2), u2);
            break;
          case IEXA_OP_POW: {
            int lg = d.un(D_LOG, u1);
            int pm1 = d.pow(u1, d.sub(u2, one));
            in.y1 = d.mul(u2, pm1);
            in.y2 = d.mul(v, lg);
            in.h11 = d.mul(d.mul(u2, d.sub(u2, one)), d.pow(u1, d.sub(u2, d.cnst(2.0))));
            in.h12 = d.mul(pm1, d.add(one, d.mul(u2, lg)));
            in.h22 = d.mul(in.y2, lg);
            break;
          }
        }
      } else if (in.kind == K_UN) {
        bool first_active = info_[nd.a].kind != K_CONSTCLASS;
        if (first_active) { // x op c
          int c = u2;
          switch (nd.op) {
            case IEXA_OP_ADD: case IEXA_OP_SUB: in.y1 = one; in.h11 = zero; break;
            case IEXA_OP_MUL: in.y1 = c; in.h11 = zero; break;
            case IEXA_OP_DIV: in.y1 = d.div(one, c); in.h11 = zero; break;
            case IEXA_OP_POW:
              in.y1 = d.mul(c, d.pow(u1, d.sub(c, one)));
              in.h11 = d.mul(d.mul(c, d.sub(c, one)), d.pow(u1, d.sub(c, d.cnst(2.0))));
              break;
          }
        } else { // c op x
          int c = u1;
          switch (nd.op) {
            case IEXA_OP_ADD: in.y1 = one; in.h11 = zero; break;
            case IEXA_OP_SUB: in.y1 = d.cnst(-1.0); in.h11 = zero; break;
            case IEXA_OP_MUL: in.y1 = c; in.h11 = zero; break;
            case IEXA_OP_DIV: in.y1 = d.neg(d.div(v, u2)); in.h11 = d.div(d.mul(d.cnst(-2.0), in.y1), u2); break;
            case IEXA_OP_POW: { int lg = d.un(D_LOG, c); in.y1 = d.mul(v, lg); in.h11 = d.mul(in.y1, lg); break; }
          }
        }
      }
    }
  }

  // ---- first order: jrpass / grpass (leaves left to right) --------------------------
  void emit1(int32_t u, int value) {
    ++g.n_occ1;
    auto it = slot1_of_.find(u);
    int32_t s;
    if (it == slot1_of_.end()) {
      s = (int32_t)g.jac_slot.size();
      slot1_of_[u] = s;
      g.jac_slot.push_back(u);
      slot1_.push_back(dag_.cnst(0.0));
    } else s = it->second;
    slot1_[s] = dag_.add(slot1_[s], value);
  }
  void jr(int t, int adj) {
    const Info &in = info_[t];
    switch (in.kind) {
      case K_VAR: emit1(g.idx_map[g.tape[t].a], adj); break;
      case K_UN: jr(in.c1, dag_.mul(adj, in.y1)); break;
      case K_BIN:
        if (order_ == 0) { jr(in.c1, dag_.mul(adj, in.y1)); jr(in.c2, dag_.mul(adj, in.y2)); }
        else { jr(in.c2, dag_.mul(adj, in.y2)); jr(in.c1, dag_.mul(adj, in.y1)); }
        break;
      default: break;
    }
  }
  void first_order() {
    if (!g.is_null) jr(n_ - 1, dag_.cnst(1.0));
    g.o1step = (int32_t)g.jac_slot.size();
  }

  // ---- second order: hrpass0 / hrpass / hdrpass --------------------------------------
  void emit2(int32_t u1, int32_t u2, int value) {
    ++g.n_occ2;
    auto key = std::make_pair(u1, u2);
    auto it = slot2_of_.find(key);
    int32_t s;
    if (it == slot2_of_.end()) {
      s = (int32_t)g.hess_slot.size();
      slot2_of_[key] = s;
      g.hess_slot.push_back(key);
      slot2_.push_back(dag_.cnst(0.0));
    } else s = it->second;
    slot2_[s] = dag_.add(slot2_[s], value);
  }
  // hrpass0: the top-level pass carries NO second-order adjoint ("adj2 == 0, no cross term"): through
  // +, -, const*subtree only the first-order adjoint is pushed down, and the FIRST nonlinear node hands its children
  // adj*h directly — not 0*y^2 + adj*h, which would turn an infinite local partial into NaN where the reference has none
  void hr0(int t, int adj) {
    const Info &in = info_[t];
    Dag &d = dag_;
    switch (in.pass) {
      case P_KEEP: hr0(in.c1, adj); return;
      case P_FLIP: hr0(in.c1, d.neg(adj)); return;
      case P_SCALE: hr0(in.c1, d.mul(adj, in.y1)); return;
      case P_ADD: case P_SUB: {
        const int a2 = in.pass == P_SUB ? d.neg(adj) : adj;
        if (order_ == 0) { hr0(in.c1, adj); hr0(in.c2, a2); } else { hr0(in.c2, a2); hr0(in.c1, adj); }
        return;
      }
      default: break;
    }
    switch (in.kind) { // linear leaves: no Hessian slot
      case K_UN: hr(in.c1, d.mul(adj, in.y1), d.mul(adj, in.h11)); break;
      case K_BIN:
        if (order_ == 0) {
          hr(in.c1, d.mul(adj, in.y1), d.mul(adj, in.h11));
          hr(in.c2, d.mul(adj, in.y2), d.mul(adj, in.h22));
          hdr(in.c1, in.c2, d.mul(adj, in.h12));
        } else {
          hr(in.c2, d.mul(adj, in.y2), d.mul(adj, in.h22));
          hr(in.c1, d.mul(adj, in.y1), d.mul(adj, in.h11));
          hdr(in.c2, in.c1, d.mul(adj, in.h12));
        }
        break;
      default: break;
    }
  }
  void hr(int t, int adj, int adj2) {
    const Info &in = info_[t];
    Dag &d = dag_;
    switch (in.kind) {
      case K_VAR: { int32_t u = g.idx_map[g.tape[t].a]; emit2(u, u, adj2); break; }
      case K_UN:
        hr(in.c1, d.mul(adj, in.y1), d.add(d.mul(adj2, d.sq(in.y1)), d.mul(adj, in.h11)));
        break;
      case K_BIN: {
        int cross = d.add(d.mul(d.mul(adj2, in.y1), in.y2), d.mul(adj, in.h12));
        if (order_ == 0) {
          hr(in.c1, d.mul(adj, in.y1), d.add(d.mul(adj2, d.sq(in.y1)), d.mul(adj, in.h11)));
          hr(in.c2, d.mul(adj, in.y2), d.add(d.mul(adj2, d.sq(in.y2)), d.mul(adj, in.h22)));
          hdr(in.c1, in.c2, cross);
        } else {
          hr(in.c2, d.mul(adj, in.y2), d.add(d.mul(adj2, d.sq(in.y2)), d.mul(adj, in.h22)));
          hr(in.c1, d.mul(adj, in.y1), d.add(d.mul(adj2, d.sq(in.y1)), d.mul(adj, in.h11)));
          hdr(in.c2, in.c1, cross);
        }
        break;
      }
      default: break;
    }
  }
  // cross pass over two subtrees: every (leaf of t1, leaf of t2) pair, children visited in the policy's order
  void hdr(int t1, int t2, int adj) {
    const Info &a = info_[t1], &b = info_[t2];
    Dag &d = dag_;
    if (a.kind == K_VAR && b.kind == K_VAR) {
      int32_t u1 = g.idx_map[g.tape[t1].a], u2 = g.idx_map[g.tape[t2].a];
      emit2(u1, u2, d.mul(adj, d.sel2(u1, u2)));
      return;
    }
    // (child, local partial) lists; a Var leaf stands for itself with partial 1
    int ca[2], ya[2], na = 0, cb[2], yb[2], nb = 0;
    auto kids = [&](const Info &n, int t, int *c, int *y, int &cnt) {
      if (n.kind == K_VAR) { c[0] = t; y[0] = -1; cnt = 1; }
      else if (n.kind == K_UN) { c[0] = n.c1; y[0] = n.y1; cnt = 1; }
      else if (n.kind == K_BIN) {
        if (order_ == 0) { c[0] = n.c1; y[0] = n.y1; c[1] = n.c2; y[1] = n.y2; }
        else { c[0] = n.c2; y[0] = n.y2; c[1] = n.c1; y[1] = n.y1; }
        cnt = 2;
      } else cnt = 0;
    };
    kids(a, t1, ca, ya, na);
    kids(b, t2, cb, yb, nb);
    for (int i = 0; i < na; ++i)
      for (int j = 0; j < nb; ++j) {
        int v = adj;
        if (ya[i] >= 0) v = d.mul(v, ya[i]);
        if (yb[j] >= 0) v = d.mul(v, yb[j]);
        hdr(ca[i], cb[j], v);
      }
  }
  void second_order() {
    if (!g.is_null) hr0(n_ - 1, dag_.w(wid_));
    g.o2step = (int32_t)g.hess_slot.size();
  }

};

// ---- DAG -> register program -----------------------------------------------------------------
inline Program schedule(const Dag &dag, const std::vector<int> &outs, size_t n_islots, std::vector<uint8_t> &x_slots) {
  Program P;
  const auto &N = dag.nodes;
  int nn = (int)N.size();
  std::vector<uint8_t> live(nn, 0);
  std::vector<int> stack(outs.begin(), outs.end());
  while (!stack.empty()) {
    int v = stack.back(); stack.pop_back();
    if (live[v]) continue;
    live[v] = 1;
    int op = N[v].op;
    if (dop_is_unary(op)) stack.push_back(N[v].a);
    else if (dop_is_binary(op)) { stack.push_back(N[v].a); stack.push_back(N[v].b); }
  }
  std::vector<int> last(nn, -1);
  for (int v = 0; v < nn; ++v) {
    if (!live[v]) continue;
    int op = N[v].op;
    if (dop_is_unary(op)) last[N[v].a] = v;
    else if (dop_is_binary(op)) { last[N[v].a] = v; last[N[v].b] = v; }
  }
  std::vector<std::vector<int>> outs_of(nn);
  for (size_t j = 0; j < outs.size(); ++j) outs_of[outs[j]].push_back((int)j);
  std::vector<int> reg(nn, -1), cidx(nn, -1);
  std::vector<int> freelist;
  int nreg = 0;
  auto alloc = [&]() { if (!freelist.empty()) { int r = freelist.back(); freelist.pop_back(); return r; } return nreg++; };
  auto operand = [&](int v) -> int32_t {
    if (N[v].op == D_CONST) {
      if (cidx[v] < 0) { cidx[v] = (int)P.cpool.size(); P.cpool.push_back(N[v].c); }
      return ~cidx[v];
    }
    return reg[v];
  };
  x_slots.assign(n_islots, 0);
  P.nout = (int32_t)outs.size();
  for (int v = 0; v < nn; ++v) {
    if (!live[v]) continue;
    int op = N[v].op;
    if (op == D_CONST) {
      for (int j : outs_of[v]) P.code.push_back(Instr{D_OUT, j, operand(v), 0});
      continue;
    }
    Instr I{op, 0, 0, 0};
    if (dop_is_unary(op)) I.a = operand(N[v].a);
    else if (dop_is_binary(op)) { I.a = operand(N[v].a); I.b = operand(N[v].b); }
    else { I.a = N[v].a; I.b = N[v].b; }
    if (op == D_LOADX) x_slots[N[v].a] = 1;
    if (op == D_W) P.uses_w = true;
    // free operand registers whose last use is here BEFORE allocating dst (dst may reuse them)
    if (dop_is_unary(op) || dop_is_binary(op)) {
      int a = N[v].a, b = dop_is_binary(op) ? N[v].b : -1;
      if (N[a].op != D_CONST && last[a] == v) freelist.push_back(reg[a]);
      if (b >= 0 && b != a && N[b].op != D_CONST && last[b] == v) freelist.push_back(reg[b]);
      ++P.n_flop_nodes;
    }
    reg[v] = alloc();
    I.dst = reg[v];
    P.code.push_back(I);
    for (int j : outs_of[v]) P.code.push_back(Instr{D_OUT, j, reg[v], 0});
    if (last[v] < 0) freelist.push_back(reg[v]); // only feeds outputs (already emitted)
  }
  P.nreg = nreg;
  return P;
}

} // namespace iexa

// plan.hpp — host-side evaluation plan: x/theta layout, SoA iterator columns, compiled
// generators, global + per-rank COO layout.  No CUDA in this header (the device engine
// lives in engine.cu; tests/hostcheck.cpp reuses this header to validate the plan
// compiler on machines without a GPU).
//
// Emission order == layout order, as fixed by build_exa_core! (src/transform.jl:771-796):
// variables / parameters / constraints / objectives are appended in call order; rows of
// generator g are o0_g + k, Jacobian slots o1_g + o1step_g*k + c, Hessian slots
// o2_g + o2step_g*k + c with all OBJECTIVE generators first.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <string>

#include "gen.hpp"

namespace iexa {

struct HostColumn {
  int64_t K = 0;
  bool is_int = false;
  bool iota = false;            // int column equal to 1..K: never stored
  // int column that is an affine pattern of its position,  v(j) = aa + ab*(j / ac) + ad*(j % ac):  never stored or
  // loaded either — index arithmetic instead of a column load followed by a dependent x load.  Covers what the
  // reference's transcription produces besides 1..K (transform.jl:27-31): restricted / shifted support indices
  // 2..T of finite-difference rows (:535-538), the (lower bound, node) pairs of orthogonal collocation (:485-505)
  // and the (upper bound, internal node) pairs of the collocation restrictions (:582-584).
  bool affine = false;
  int64_t aa = 0, ab = 0, ac = 1, ad = 0;
  std::vector<int32_t> ivals;   // int column (narrowed; x/theta indices fit in int32)
  int32_t vmin = 0, vmax = 0;   // min / max of a stored int column (index-range queries over the whole column)
  std::vector<double> fvals;
  // GENERATED fp column (iexa_itr_generated): described by a closed form and produced ON the device by a kernel — no
  // K-long host array, no upload.  The formulas restate numpy / InfiniteOpt arithmetic operation by operation (each
  // product and sum rounded once, no FMA), so the device values are bit-identical to the host path's:
  //   LINSPACE      v(j) = a + j*step, step = (b-a)/(n-1), v(n-1) = b          (supports of an independent parameter, transform.jl:22-24)
  //   LINSPACE_MID  the same public grid interleaved with the interval midpoints: v(2i) = pub(i), v(2i+1) = 0.5*(pub(i)+pub(i+1)), K = 2n-1
  //                 (one internal Lobatto node per interval: OrthogonalCollocation(3), transform.jl:22)
  //   TRAPEZOID     c(j) = (s(j+1)-s(j))/2 [j < K-1] + (s(j)-s(j-1))/2 [j > 0] of column gen_src   (InfiniteOpt's default integral, transform.jl:625-626)
  //   CONST         v(j) = a
  int32_t gen_kind = 0;         // 0: stored column
  int32_t gen_src = -1;         // TRAPEZOID: index into Plan::columns
  int64_t gen_n = 0;
  double gen_a = 0, gen_b = 0;
  double gen_value(int64_t j, const std::vector<HostColumn> &cols) const {
    auto pub = [&](int64_t i) { return i == gen_n - 1 ? gen_b : gen_a + (double)i * ((gen_b - gen_a) / (double)(gen_n - 1)); };
    switch (gen_kind) {
      case 1: return gen_n == 1 ? gen_a : pub(j);
      case 2: return (j & 1) ? 0.5 * (pub(j >> 1) + pub((j >> 1) + 1)) : pub(j >> 1);
      case 3: {
        const HostColumn &s = cols[gen_src];
        double c = 0.0;
        if (j < K - 1) c = (s.fp(j + 1, cols) - s.fp(j, cols)) / 2;
        if (j > 0) c = c + (s.fp(j, cols) - s.fp(j - 1, cols)) / 2;
        return c;
      }
      case 4: return gen_a;
      default: return 0.0;
    }
  }
  double fp(int64_t j, const std::vector<HostColumn> &cols) const { return gen_kind ? gen_value(j, cols) : fvals[j]; }
  int64_t ival(int64_t j) const { return affine ? aa + ab * (j / ac) + ad * (j % ac) : (int64_t)ivals[j]; }
  // v(j) = a + b*(j/c) + d*(j%c) for some period c <= 8?
  bool detect_affine(const int64_t *v, int64_t n) {
    if (getenv("IEXA_NO_AFFINE")) return false;
    if (n <= 0) { affine = true; aa = 1; ab = 1; ac = 1; ad = 0; return true; }
    for (int64_t c = 1; c <= 8 && (c == 1 || c < n); ++c) {
      const int64_t a = v[0], d = (c > 1 && n > 1) ? v[1] - v[0] : 0, b = n > c ? v[c] - v[0] : 0;
      bool ok = true;
      for (int64_t k = 0; k < n && ok; ++k) ok = v[k] == a + b * (k / c) + d * (k % c);
      if (ok) { affine = true; aa = a; ab = b; ac = c; ad = d; return true; }
    }
    return false;
  }
};

struct ColRef {
  int32_t col; // index into Plan::columns
  int64_t div, mod; // value(k) = column[(k / div) % mod]
};

struct Iterator {
  int64_t K = 1;
  std::vector<ColRef> int_cols, fp_cols;
};

struct Generator {
  bool is_obj = false;
  int32_t itr = 0;
  int64_t K = 1;
  double lcon = 0, ucon = 0;
  GenCompiled c;
  // global layout (0-based)
  int64_t o0 = 0, o1 = 0, o2 = 0, og = 0;
  // this rank's support range and local layout
  int64_t k0 = 0, k1 = 0;
  int64_t l0 = 0, l1 = 0, l2 = 0;
  // objective generators: first-order slots that are the ONLY writer of their gradient entries
  // (plain store instead of an atomic, and no zero-fill of that range) — Plan::analyse_grad
  std::vector<uint8_t> grad_direct;
};

// Generators over the SAME iterator are fused into one program: one thread evaluates every
// member at its support point, so shared loads (x, theta, columns) and shared sub-expressions
// (sin/cos of the same state) are done once.  Layout is untouched: each member still owns its
// rows / slots.  (The reference stack evaluates every generator in its own kernel launch.)
struct Group {
  bool is_obj = false;
  int32_t itr = 0;
  int64_t K = 1;
  int64_t k0 = 0, k1 = 0;
  std::vector<int32_t> members; // indices into Plan::objs or Plan::cons
  SlotCtx ctx;
  // val, d1, d2 over all members, then the matrix-free products jv (one output per member), jtv and hv (one output per
  // touched index slot of the group: contributions of all members are summed in the program)
  // [6] (GPROG_ALL, constraint groups): value + first + second order of every member in ONE program — the fused
  // cons! + jac_coord! + hess_coord! evaluation (iexa_eval3): x / theta / columns loaded once, sin / cos of a state once
  Program prog[7];
  std::vector<std::pair<int32_t, int32_t>> outmap[7]; // program output j -> (member position, slot); jtv / hv: (0, index slot); [6]: (3*member + {0 val, 1 d1, 2 d2}, slot)
  std::vector<std::vector<int32_t>> jac_slot;         // per member: group index slot of each first-order slot
  std::vector<uint8_t> x_slots[7];
  // scatter products (jtprod! / hprod!): phase of the group (0: launched first, may store single-writer outputs directly;
  // 1: launched after phase 0, atomics only) and, per output, 1 = plain store — Plan::analyse_scatter
  int scat_phase[2] = {1, 1};
  std::vector<uint8_t> scat_direct[2];   // per output: 0 atomic, 1 plain store, 2 accumulate onto the store of the same thread (rider)
  // RIDER: a group with the SAME number of supports as an accepted phase-0 group whose single-writer outputs address the same entries
  // through the same thread (index c0 + s*k with equal c0, s) is evaluated by that group's threads right after its body —
  // `out[i] += v` by the thread that just stored out[i]: no atomic, no zero-fill, x / v loads hit the L1 the primary
  // filled.  hprod! of config 3: the fused ODE rows ride on the objective group (both walk the time supports).
  int scat_rider_of[2] = {-1, -1};
  size_t dag_nodes = 0;
  // Shape class (build_groups(class_mode = true)): ONE member program shared by many generators of
  // identical shape — same tape structure, columns and index-term structure; they differ only in
  // literal constants, index bases and output offsets, which become per-instance table entries.
  // This is what turns the reference's "one generator per bus / branch constraint" launch storm
  // (ESCAPE34/opf.jl:150-283) into a handful of kernel bodies with an instance axis.
  // FUSED classes: classes with the same number of instances over the same iterator whose q-th instances read the same
  // variables (the P / Q flow rows and the thermal limit of ONE branch: vm[f], vm[t], va[f], va[t], p, q) become the
  // members of ONE class group — instance q evaluates the q-th generator of every member class in one program, so the
  // shared loads and the sin / cos of the angle difference are done once per branch and scenario instead of once per row.
  bool is_class = false;
  std::vector<int32_t> inst_gens;   // [instance q][member j] -> generator index, flattened q * members.size() + j
  std::vector<int64_t> inst_base;   // [instance q][group index slot] -> index base of that instance, flattened
  std::vector<int32_t> cpar_nodes;  // per-instance constant c: tape node id ...
  std::vector<int32_t> cpar_member; // ... of member cpar_member[c]
  size_t n_inst() const { return members.empty() ? 0 : inst_gens.size() / members.size(); }
  int32_t inst_gen(size_t q, size_t j) const { return inst_gens[q * members.size() + j]; }
};

struct Plan {
  bool minimize = true;
  bool finalized = false;
  // iexa_set_option: slot-order policy of the symbolic passes (gen.hpp: GenCompiler::order_) and IEEE-strict folding
  // (dag.hpp: Dag::strict); both must be chosen before the first generator is added
  int opt_slot_order = 0;
  bool opt_strict = true;   // default: IEEE-strict (measured cost on B200: +1.5 % per eval on config 3, +2.5 % on the 118-bus OPF)
  std::vector<double> x0, lvar, uvar, theta;
  std::vector<HostColumn> columns;
  std::vector<Iterator> itrs;
  std::vector<Generator> objs, cons;
  std::vector<Group> groups; // built by layout(); objective groups first
  int64_t nvar = 0, npar = 0, ncon = 0, nnzj = 0, nnzh = 0, nnzg = 0;
  int64_t loc_ncon = 0, loc_nnzj = 0, loc_nnzh = 0;
  int32_t rank = 0, world = 1, device = -1;
  std::vector<double> y0; // multipliers start: allocated on the first iexa_set_vector(5); all zeros until then
  // lcon / ucon are constants per generator (transform.jl:396-411): filled on demand by fill_con_bounds()

  Plan() {
    itrs.emplace_back(); // iterator 0 is the empty iterator [(;)] (transform.jl:440, :614)
  }

  // x0 / lvar / uvar are kept per add_var block — a NULL array is the default (0, -inf, +inf) and costs nothing — and the dense
  // vectors are materialised only when somebody asks for them (iexa_get_vector, iexa_patch_var, iexa_set_vector): the kernels never
  // read them, and at 10^6 supports the three 352 MB host vectors were 0.9 s of a 1.0 s plan build.
  struct VarBlock { int64_t n; std::vector<double> s, l, u; };
  std::vector<VarBlock> vblocks;
  bool vars_dense = false;
  int64_t add_var(int64_t n, const double *s, const double *l, const double *u) {
    int64_t off = nvar;
    if (vars_dense) { // after a patch: keep the dense vectors current
      const double inf = std::numeric_limits<double>::infinity();
      if (s) x0.insert(x0.end(), s, s + n); else x0.resize(x0.size() + n, 0.0);
      if (l) lvar.insert(lvar.end(), l, l + n); else lvar.resize(lvar.size() + n, -inf);
      if (u) uvar.insert(uvar.end(), u, u + n); else uvar.resize(uvar.size() + n, inf);
    } else {
      vblocks.emplace_back();
      VarBlock &b = vblocks.back();
      b.n = n;
      if (s) b.s.assign(s, s + n);
      if (l) b.l.assign(l, l + n);
      if (u) b.u.assign(u, u + n);
    }
    nvar += n;
    return off;
  }
  void materialise_vars() {
    if (vars_dense) return;
    const double inf = std::numeric_limits<double>::infinity();
    x0.clear(); lvar.clear(); uvar.clear();
    x0.reserve(nvar); lvar.reserve(nvar); uvar.reserve(nvar);
    for (VarBlock &b : vblocks) {
      if (b.s.empty()) x0.resize(x0.size() + b.n, 0.0); else x0.insert(x0.end(), b.s.begin(), b.s.end());
      if (b.l.empty()) lvar.resize(lvar.size() + b.n, -inf); else lvar.insert(lvar.end(), b.l.begin(), b.l.end());
      if (b.u.empty()) uvar.resize(uvar.size() + b.n, inf); else uvar.insert(uvar.end(), b.u.begin(), b.u.end());
    }
    vblocks.clear(); vblocks.shrink_to_fit();
    vars_dense = true;
  }
  int64_t add_par(int64_t n, const double *v) {
    int64_t off = npar;
    theta.insert(theta.end(), v, v + n);
    npar += n;
    return off;
  }

  // dense lcon (upper == false) / ucon of the global row numbering, straight into the caller's buffer
  void fill_con_bounds(bool upper, double *out) const {
    for (const Generator &g : cons) std::fill(out + g.o0, out + g.o0 + g.K, upper ? g.ucon : g.lcon);
  }

  int32_t itr_base(int64_t K, int32_t n_int, const int64_t *const *ic, int32_t n_fp,
                   const double *const *fc) {
    if (K < 0) throw std::invalid_argument("iterator: negative length");
    Iterator it;
    it.K = K;
    for (int32_t j = 0; j < n_int; ++j) {
      HostColumn c;
      c.K = K; c.is_int = true; c.iota = true;
      for (int64_t k = 0; k < K; ++k)
        if (ic[j][k] != k + 1) { c.iota = false; break; }
      if (c.iota) { c.affine = true; c.aa = 1; c.ab = 1; c.ac = 1; c.ad = 0; }
      else c.detect_affine(ic[j], K);
      if (!c.affine) {
        c.ivals.resize(K);
        for (int64_t k = 0; k < K; ++k) {
          int64_t v = ic[j][k];
          if (v < INT32_MIN || v > INT32_MAX) throw std::invalid_argument("iterator: integer field exceeds int32");
          c.ivals[k] = (int32_t)v;
        }
        if (K > 0) { c.vmin = *std::min_element(c.ivals.begin(), c.ivals.end()); c.vmax = *std::max_element(c.ivals.begin(), c.ivals.end()); }
      }
      it.int_cols.push_back(ColRef{(int32_t)columns.size(), 1, K > 0 ? K : 1});
      columns.push_back(std::move(c));
    }
    for (int32_t j = 0; j < n_fp; ++j) {
      HostColumn c;
      c.K = K; c.is_int = false;
      c.fvals.assign(fc[j], fc[j] + K);
      it.fp_cols.push_back(ColRef{(int32_t)columns.size(), 1, K > 0 ? K : 1});
      columns.push_back(std::move(c));
    }
    itrs.push_back(std::move(it));
    return (int32_t)itrs.size() - 1;
  }

  // iterator whose columns are (partly) GENERATED: int_cols[j] == nullptr is 1..K; gens[j].kind != 0 describes fp column j
  // by a closed form (HostColumn::gen_*), kind == 0 takes fp_cols[j] as data
  int32_t itr_generated(int64_t K, int32_t n_int, const int64_t *const *ic, int32_t n_fp, const iexa_colgen *gens,
                        const double *const *fc) {
    if (K < 0) throw std::invalid_argument("iterator: negative length");
    // integer columns: data (through the stored path: affine detection etc.) or iota
    std::vector<std::vector<int64_t>> iota_store;
    std::vector<const int64_t *> icp(n_int, nullptr);
    const size_t col0 = columns.size();
    Iterator it;
    it.K = K;
    for (int32_t j = 0; j < n_int; ++j) {
      HostColumn c;
      c.K = K; c.is_int = true;
      if (!ic || !ic[j]) { c.iota = true; c.affine = true; c.aa = 1; c.ab = 1; c.ac = 1; c.ad = 0; }
      else {
        c.iota = true;
        for (int64_t k = 0; k < K; ++k) if (ic[j][k] != k + 1) { c.iota = false; break; }
        if (c.iota) { c.affine = true; c.aa = 1; c.ab = 1; c.ac = 1; c.ad = 0; }
        else c.detect_affine(ic[j], K);
        if (!c.affine) {
          c.ivals.resize(K);
          for (int64_t k = 0; k < K; ++k) {
            if (ic[j][k] < INT32_MIN || ic[j][k] > INT32_MAX) throw std::invalid_argument("iterator: integer field exceeds int32");
            c.ivals[k] = (int32_t)ic[j][k];
          }
          if (K > 0) { c.vmin = *std::min_element(c.ivals.begin(), c.ivals.end()); c.vmax = *std::max_element(c.ivals.begin(), c.ivals.end()); }
        }
      }
      it.int_cols.push_back(ColRef{(int32_t)columns.size(), 1, K > 0 ? K : 1});
      columns.push_back(std::move(c));
    }
    const size_t fcol0 = columns.size();
    for (int32_t j = 0; j < n_fp; ++j) {
      HostColumn c;
      c.K = K; c.is_int = false;
      const iexa_colgen &gs = gens[j];
      if (gs.kind == 0) {
        if (!fc || !fc[j]) throw std::invalid_argument("iterator: fp column without data or generator");
        c.fvals.assign(fc[j], fc[j] + K);
      } else {
        c.gen_kind = gs.kind; c.gen_n = gs.n; c.gen_a = gs.a; c.gen_b = gs.b;
        if (gs.kind == 1 && (gs.n != K || K < 1)) throw std::invalid_argument("generated column: LINSPACE needs n == K >= 1");
        if (gs.kind == 2 && (gs.n < 2 || 2 * gs.n - 1 != K)) throw std::invalid_argument("generated column: LINSPACE_MID needs K == 2n-1, n >= 2");
        if (gs.kind == 3) {
          if (gs.src < 0 || gs.src >= j) throw std::invalid_argument("generated column: TRAPEZOID needs an earlier fp column of the same iterator");
          c.gen_src = (int32_t)(fcol0 + gs.src);
        }
        if (gs.kind < 0 || gs.kind > 4) throw std::invalid_argument("generated column: unknown kind");
      }
      it.fp_cols.push_back(ColRef{(int32_t)columns.size(), 1, K > 0 ? K : 1});
      columns.push_back(std::move(c));
    }
    (void)col0;
    itrs.push_back(std::move(it));
    return (int32_t)itrs.size() - 1;
  }

  // Parameter functions (transform.jl:161-183: a block of theta holding f(supports) at every support combination) as
  // TAPES: the engine evaluates them on the device at finalize with the same register programs / interpreter kernel as
  // a constraint's value — no O(K) closure calls on the host, no upload of the block.  Leaves: literals, fp fields of the
  // iterator, theta (finite parameters added earlier).
  std::vector<Generator> pfuncs;
  int64_t add_par_function(const iexa_node *nodes, int32_t n, const iexa_index *idx, int32_t n_idx, int32_t itr) {
    for (int32_t i = 0; i < n; ++i) if (nodes[i].op == IEXA_OP_VAR) throw std::invalid_argument("parameter function: variables are not allowed");
    Generator g = make_gen(nodes, n, idx, n_idx, itr, false);
    const int64_t off = npar;
    g.o0 = g.l0 = off; g.k0 = 0; g.k1 = g.K;
    theta.resize(theta.size() + g.K, 0.0);
    npar += g.K;
    pfuncs.push_back(std::move(g));
    return off;
  }

  int32_t itr_product(int32_t n, const int32_t *ids) {
    Iterator it;
    it.K = 1;
    for (int32_t f = 0; f < n; ++f) {
      if (ids[f] < 0 || ids[f] >= (int32_t)itrs.size()) throw std::invalid_argument("product: bad iterator id");
      const Iterator &s = itrs[ids[f]];
      for (auto c : s.int_cols) { c.div *= it.K; it.int_cols.push_back(c); }
      for (auto c : s.fp_cols) { c.div *= it.K; it.fp_cols.push_back(c); }
      it.K *= s.K;
    }
    itrs.push_back(std::move(it));
    return (int32_t)itrs.size() - 1;
  }

  Generator make_gen(const iexa_node *nodes, int32_t n, const iexa_index *idx, int32_t n_idx,
                     int32_t itr, bool is_obj) {
    if (finalized) throw std::logic_error("plan already finalized");
    if (itr < 0 || itr >= (int32_t)itrs.size()) throw std::invalid_argument("bad iterator id");
    const Iterator &it = itrs[itr];
    GenCompiler gc(nodes, n, idx, n_idx, (int32_t)it.int_cols.size(), (int32_t)it.fp_cols.size());
    gc.is_obj_ = is_obj;
    gc.set_options(opt_slot_order == 2 ? 0 : opt_slot_order, opt_strict);
    std::vector<int32_t> perm;
    bool sorted = false;
    if (opt_slot_order == 2 && !is_obj) {
      // a dry symbolic pass gives the policy-order slots; their static column order (if there is one) becomes the slot order
      GenCompiler probe(nodes, n, idx, n_idx, (int32_t)it.int_cols.size(), (int32_t)it.fp_cols.size());
      probe.set_options(0, opt_strict);
      probe.differentiate();
      sorted = jac_row_sort(it, probe.ctx(), probe.g.jac_slot, perm);
      if (sorted) gc.jac_perm = &perm;
    }
    gc.compile();
    Generator g;
    g.itr = itr;
    g.K = it.K;
    g.c = std::move(gc.g);
    if (sorted) { g.c.jac_perm = perm; g.c.jac_row_sorted = true; }
    return g;
  }

  int64_t add_con(const iexa_node *nodes, int32_t n, const iexa_index *idx, int32_t n_idx,
                  int32_t itr, double lc, double uc) {
    Generator g = make_gen(nodes, n, idx, n_idx, itr, false);
    g.is_obj = false; g.lcon = lc; g.ucon = uc;
    g.o0 = ncon;
    ncon += g.K;

    cons.push_back(std::move(g));
    return cons.back().o0;
  }
  void add_obj(const iexa_node *nodes, int32_t n, const iexa_index *idx, int32_t n_idx, int32_t itr) {
    Generator g = make_gen(nodes, n, idx, n_idx, itr, true);
    g.is_obj = true;
    objs.push_back(std::move(g));
  }

  // column accessors for a generator (host side: structure checks, hostcheck, sharding analysis)
  int64_t int_col_value(const Generator &g, int32_t slot, int64_t k) const {
    const ColRef &r = itrs[g.itr].int_cols[g.c.int_cols[slot]];
    int64_t j = (k / r.div) % r.mod;
    const HostColumn &c = columns[r.col];
    return c.ival(j);
  }
  double fp_col_value(const Generator &g, int32_t slot, int64_t k) const {
    const ColRef &r = itrs[g.itr].fp_cols[g.c.fp_cols[slot]];
    return columns[r.col].fp((k / r.div) % r.mod, columns);
  }
  int64_t index_value(const Generator &g, int32_t islot, int64_t k) const {
    const IndexExpr &e = g.c.uidx[islot];
    int64_t v = e.base;
    for (auto &t : e.terms) v += t.second * int_col_value(g, t.first, k);
    return v;
  }

  // contiguous support blocks; generators shorter than the world (length-1 generators: point constraints,
  // non-measure objective terms) go one support per rank starting at rank 0
  static void shard_range(int64_t K, int64_t r, int64_t w, int64_t &k0, int64_t &k1) {
    if (K < w) { k0 = std::min<int64_t>(r, K); k1 = std::min<int64_t>(r + 1, K); return; }
    k0 = (K * r) / w;
    k1 = (K * (r + 1)) / w;
  }

  void layout(int32_t rank_, int32_t world_) {
    if (world_ < 1 || rank_ < 0 || rank_ >= world_) throw std::invalid_argument("bad rank/world");
    check_index_bounds();
    rank = rank_; world = world_;
    nnzj = nnzh = nnzg = 0;
    loc_ncon = loc_nnzj = loc_nnzh = 0;
    auto range = [&](Generator &g) { shard_range(g.K, rank, world, g.k0, g.k1); };
    for (auto &g : objs) {
      range(g);
      g.og = nnzg; nnzg += g.K * g.c.o1step;
      g.o2 = nnzh; nnzh += g.K * g.c.o2step;
      g.l2 = loc_nnzh; loc_nnzh += (g.k1 - g.k0) * g.c.o2step;
    }
    for (auto &g : cons) {
      range(g);
      g.o1 = nnzj; nnzj += g.K * g.c.o1step;
      g.o2 = nnzh; nnzh += g.K * g.c.o2step;
      g.l0 = loc_ncon; loc_ncon += (g.k1 - g.k0);
      g.l1 = loc_nnzj; loc_nnzj += (g.k1 - g.k0) * g.c.o1step;
      g.l2 = loc_nnzh; loc_nnzh += (g.k1 - g.k0) * g.c.o2step;
    }
    shard_pfuncs();
    build_groups();
    analyse_grad();
    finalized = true;
  }

  // shape key of a generator: everything except literal values, index bases and bounds
  static std::string shape_key(const Generator &g) {
    std::string k;
    auto put = [&](int64_t v) { k.append(reinterpret_cast<const char *>(&v), sizeof v); };
    put(g.itr); put((int64_t)g.c.tape.size());
    for (const iexa_node &n : g.c.tape) { put(n.op); if (n.op != IEXA_OP_CONST && n.op != IEXA_OP_VAR && n.op != IEXA_OP_PAR) { put(n.a); put(n.b); } }
    // leaves: which index SLOT each VAR/PAR leaf uses (identity pattern) and each slot's term structure
    for (const iexa_node &n : g.c.tape) if (n.op == IEXA_OP_VAR || n.op == IEXA_OP_PAR) put(g.c.idx_map[n.a]);
    for (const IndexExpr &e : g.c.uidx) { put((int64_t)e.terms.size()); for (auto &t : e.terms) { put(t.first); put(t.second); } }
    for (int32_t c : g.c.int_cols) put(c);
    for (int32_t c : g.c.fp_cols) put(c);
    for (int32_t s : g.c.jac_slot) put(s);
    for (int32_t s : g.c.jac_perm) put(s);
    for (auto &pr : g.c.hess_slot) { put(pr.first); put(pr.second); }
    return k;
  }

  // Fuse generators that share (kind, iterator); a group is closed when it grows past the caps.
  // class_mode: generators of identical SHAPE (>= min_inst of them) become one class group first.
  void build_groups(bool class_mode = false, size_t min_inst = 4, size_t max_dag_nodes = 6000, int max_slots = 96) {
    if (const char *e = getenv("IEXA_GROUP_MAX_SLOTS")) max_slots = atoi(e);       // tuning knobs
    if (const char *e = getenv("IEXA_GROUP_MAX_NODES")) max_dag_nodes = (size_t)atoll(e);
    groups.clear();
    for (int pass = 0; pass < 2; ++pass) {
      std::vector<Generator> &gens = pass == 0 ? objs : cons;
      std::vector<uint8_t> taken(gens.size(), 0);
      if (class_mode) {
        std::map<std::string, std::vector<int32_t>> classes;
        std::vector<std::string> order;
        for (size_t gi = 0; gi < gens.size(); ++gi) {
          std::string key = shape_key(gens[gi]);
          if (!classes.count(key)) order.push_back(key);
          classes[key].push_back((int32_t)gi);
        }
        // super-classes: lists of classes whose q-th instances are fused (see Group::is_class)
        std::vector<std::vector<const std::vector<int32_t> *>> supers;
        const bool fuse = !getenv("IEXA_NO_CLASS_FUSION");
        for (const std::string &key : order) {
          const std::vector<int32_t> &inst = classes[key];
          if (inst.size() < min_inst) continue;
          bool placed = false;
          for (size_t si = supers.size(); fuse && si-- > 0 && !placed;) {
            auto &S = supers[si];
            const Generator &a0 = gens[(*S[0])[0]], &b0 = gens[inst[0]];
            if (a0.itr != b0.itr || S[0]->size() != inst.size()) continue;
            int s1 = 0, s2 = 0;
            size_t nodes = 0;
            for (auto *c : S) { const Generator &g = gens[(*c)[0]]; s1 += g.c.o1step | 1; s2 += g.c.o2step | 1; nodes += g.c.tape.size(); }
            if (s1 + (b0.c.o1step | 1) > max_slots || s2 + (b0.c.o2step | 1) > max_slots || 12 * (nodes + b0.c.tape.size()) > max_dag_nodes) continue;
            // the q-th instances must share variables the same way for EVERY q: every pair of index slots that coincides
            // in instance 0 (same terms, same base) coincides in all instances; and at least one pair does
            bool shares = false, ok = true;
            for (auto *c : S) {
              const Generator &ga = gens[(*c)[0]];
              for (size_t u = 0; u < ga.c.uidx.size() && ok; ++u)
                for (size_t w = 0; w < b0.c.uidx.size() && ok; ++w) {
                  if (!(ga.c.uidx[u] == b0.c.uidx[w])) continue;
                  if (ga.c.int_cols.size() != b0.c.int_cols.size() || ga.c.int_cols != b0.c.int_cols) { ok = false; break; } // term column slots must mean the same columns
                  shares = true;
                  for (size_t q = 1; q < inst.size() && ok; ++q)
                    ok = gens[(*c)[q]].c.uidx[u].base == gens[inst[q]].c.uidx[w].base;
                }
            }
            if (ok && shares) { S.push_back(&inst); placed = true; }
          }
          if (!placed) supers.push_back({&inst});
        }
        for (auto &S : supers) {
          const size_t m = S.size(), ninst = S[0]->size();
          const Generator &g0 = gens[(*S[0])[0]];
          const Iterator &it = itrs[g0.itr];
          groups.emplace_back();
          Group &G = groups.back();
          G.is_obj = pass == 0; G.itr = g0.itr; G.K = g0.K; G.k0 = g0.k0; G.k1 = g0.k1;
          G.is_class = true;
          G.inst_gens.resize(ninst * m);
          for (size_t q = 0; q < ninst; ++q) for (size_t j = 0; j < m; ++j) G.inst_gens[q * m + j] = (*S[j])[q];
          Dag dag;
          std::vector<MemberSlots> mslots;
          std::vector<int> o0;
          std::vector<std::vector<int>> o12(2);
          std::vector<std::vector<int32_t>> own2grp(m);
          std::vector<std::vector<int32_t>> cpars(m);
          for (size_t j = 0; j < m; ++j) {
            const std::vector<int32_t> &inst = *S[j];
            const Generator &gj = gens[inst[0]];
            G.members.push_back(inst[0]);
            // literals that differ between instances become parameters
            std::vector<int32_t> &cpar = cpars[j];
            cpar.assign(gj.c.tape.size(), -1);
            for (size_t n = 0; n < gj.c.tape.size(); ++n) {
              if (gj.c.tape[n].op != IEXA_OP_CONST) continue;
              bool same = true;
              for (int32_t gi : inst) if (std::memcmp(&gens[gi].c.tape[n].c, &gj.c.tape[n].c, 8) != 0) { same = false; break; }
              if (!same) { cpar[n] = (int32_t)G.cpar_nodes.size(); G.cpar_nodes.push_back((int32_t)n); G.cpar_member.push_back((int32_t)j); }
            }
            GenCompiler gc(gj.c.tape.data(), (int32_t)gj.c.tape.size(), gj.c.raw_idx.data(), (int32_t)gj.c.raw_idx.size(),
                           (int32_t)it.int_cols.size(), (int32_t)it.fp_cols.size(), G.ctx, dag, (int)j);
            gc.cpar_of_node = &cpar;
            gc.set_options(opt_slot_order == 2 ? 0 : opt_slot_order, opt_strict);
            if (!gj.c.jac_perm.empty()) gc.jac_perm = &gj.c.jac_perm;
            gc.differentiate();
            // a parameter that happens to be 0/1 in instance 0 must not have been folded: slot counts must match
            if ((int)gc.slot1().size() != gj.c.o1step || (int)gc.slot2().size() != gj.c.o2step)
              throw std::logic_error("shape class: sparsity of the parametrised program differs from its instances");
            own2grp[j].assign(gj.c.uidx.size(), -1);
            for (size_t i = 0; i < gj.c.idx_map.size(); ++i) own2grp[j][gj.c.idx_map[i]] = gc.g.idx_map[i];
            G.jac_slot.push_back(gc.g.jac_slot);
            o0.push_back(gc.val_root());
            G.outmap[0].push_back({(int32_t)j, 0});
            for (size_t c = 0; c < gc.slot1().size(); ++c) { o12[0].push_back(gc.slot1()[c]); G.outmap[1].push_back({(int32_t)j, (int32_t)c}); }
            for (size_t c = 0; c < gc.slot2().size(); ++c) { o12[1].push_back(gc.slot2()[c]); G.outmap[2].push_back({(int32_t)j, (int32_t)c}); }
            MemberSlots ms;
            ms.s1 = gc.slot1(); ms.s2 = gc.slot2(); ms.jac_slot = gc.g.jac_slot; ms.hess_slot = gc.g.hess_slot; ms.wid = (int)j; ms.val = gc.val_root();
            mslots.push_back(std::move(ms));
          }
          const size_t nis = G.ctx.uidx.size();
          G.inst_base.assign(ninst * nis, 0);
          for (size_t q = 0; q < ninst; ++q)
            for (size_t j = 0; j < m; ++j) {
              const Generator &gq = gens[(*S[j])[q]];
              for (size_t u = 0; u < own2grp[j].size(); ++u)
                if (own2grp[j][u] >= 0) G.inst_base[q * nis + own2grp[j][u]] = gq.c.uidx[u].base;
            }
          G.prog[0] = schedule(dag, o0, nis, G.x_slots[0]);
          G.prog[1] = schedule(dag, o12[0], nis, G.x_slots[1]);
          G.prog[2] = schedule(dag, o12[1], nis, G.x_slots[2]);
          finish_products(G, dag, mslots);
          G.dag_nodes = dag.nodes.size();
          for (auto *c : S) for (int32_t gi : *c) taken[gi] = 1;
        }
      }
      std::map<int32_t, int> open; // itr -> group index
      std::vector<std::unique_ptr<Dag>> dags;
      std::vector<std::vector<int>> outs[3];
      std::vector<int> slots1, slots2;
      std::vector<std::vector<MemberSlots>> mslots; // per local dag: the members' slot nodes (for the product programs)
      std::vector<int> gid_of_local; // local dag index -> group index
      for (size_t gi = 0; gi < gens.size(); ++gi) {
        if (taken[gi]) continue;
        Generator &g = gens[gi];
        const Iterator &it = itrs[g.itr];
        int li = -1;
        auto f = open.find(g.itr);
        if (f != open.end()) {
          int cand = f->second;
          int s1 = g.c.o1step > 1 ? (g.c.o1step | 1) : 0, s2 = g.c.o2step > 1 ? (g.c.o2step | 1) : 0;
          if (dags[cand]->nodes.size() < max_dag_nodes && slots1[cand] + s1 <= max_slots && slots2[cand] + s2 <= max_slots)
            li = cand;
        }
        if (li < 0) {
          li = (int)dags.size();
          groups.emplace_back();
          Group &G = groups.back();
          G.is_obj = pass == 0; G.itr = g.itr; G.K = g.K; G.k0 = g.k0; G.k1 = g.k1;
          dags.emplace_back(new Dag());
          for (auto &o : outs) o.emplace_back();
          slots1.push_back(0); slots2.push_back(0);
          mslots.emplace_back();
          gid_of_local.push_back((int)groups.size() - 1);
          open[g.itr] = li;
        }
        Group &G = groups[gid_of_local[li]];
        int mpos = (int)G.members.size();
        GenCompiler gc(g.c.tape.data(), (int32_t)g.c.tape.size(), g.c.raw_idx.data(), (int32_t)g.c.raw_idx.size(),
                       (int32_t)it.int_cols.size(), (int32_t)it.fp_cols.size(), G.ctx, *dags[li], mpos);
        gc.set_options(opt_slot_order == 2 ? 0 : opt_slot_order, opt_strict);
        if (!g.c.jac_perm.empty()) gc.jac_perm = &g.c.jac_perm;
        gc.differentiate();
        G.members.push_back((int32_t)gi);
        G.jac_slot.push_back(gc.g.jac_slot);
        outs[0][li].push_back(gc.val_root());
        G.outmap[0].push_back({mpos, 0});
        for (size_t c = 0; c < gc.slot1().size(); ++c) { outs[1][li].push_back(gc.slot1()[c]); G.outmap[1].push_back({mpos, (int32_t)c}); }
        for (size_t c = 0; c < gc.slot2().size(); ++c) { outs[2][li].push_back(gc.slot2()[c]); G.outmap[2].push_back({mpos, (int32_t)c}); }
        slots1[li] += g.c.o1step > 1 ? (g.c.o1step | 1) : 0;
        slots2[li] += g.c.o2step > 1 ? (g.c.o2step | 1) : 0;
        MemberSlots ms;
        ms.s1 = gc.slot1(); ms.s2 = gc.slot2(); ms.jac_slot = gc.g.jac_slot; ms.hess_slot = gc.g.hess_slot; ms.wid = mpos; ms.val = gc.val_root();
        mslots[li].push_back(std::move(ms));
      }
      for (size_t li = 0; li < dags.size(); ++li) {
        Group &G = groups[gid_of_local[li]];
        for (int p = 0; p < 3; ++p) G.prog[p] = schedule(*dags[li], outs[p][li], G.ctx.uidx.size(), G.x_slots[p]);
        finish_products(G, *dags[li], mslots[li]);
        G.dag_nodes = dags[li]->nodes.size();
      }
    }
    class_mode_ = class_mode;
    analyse_scatter();
  }
  // jv / jtv / hv programs of a group from its members' slot nodes (gen.hpp: build_products)
  static void finish_products(Group &G, Dag &dag, const std::vector<MemberSlots> &M) {
    ProductOuts po = build_products(dag, M, G.is_obj);
    const size_t nis = G.ctx.uidx.size();
    for (int p = 3; p < 6; ++p) { G.outmap[p].clear(); G.prog[p] = Program(); }
    if (!G.is_obj) {
      G.prog[3] = schedule(dag, po.jv, nis, G.x_slots[3]);
      for (size_t m = 0; m < po.jv.size(); ++m) G.outmap[3].push_back({(int32_t)m, 0});
      G.prog[4] = schedule(dag, po.jtv, nis, G.x_slots[4]);
      for (int32_t u : po.jtv_slot) G.outmap[4].push_back({0, u});
    }
    G.prog[5] = schedule(dag, po.hv, nis, G.x_slots[5]);
    for (int32_t u : po.hv_slot) G.outmap[5].push_back({0, u});
    // fused value + first + second order (iexa_eval3)
    G.outmap[6].clear(); G.prog[6] = Program();
    if (!G.is_obj) {
      std::vector<int> all;
      for (size_t m = 0; m < M.size(); ++m) {
        all.push_back(M[m].val); G.outmap[6].push_back({(int32_t)(3 * m), 0});
        for (size_t c = 0; c < M[m].s1.size(); ++c) { all.push_back(M[m].s1[c]); G.outmap[6].push_back({(int32_t)(3 * m + 1), (int32_t)c}); }
        for (size_t c = 0; c < M[m].s2.size(); ++c) { all.push_back(M[m].s2[c]); G.outmap[6].push_back({(int32_t)(3 * m + 2), (int32_t)c}); }
      }
      G.prog[6] = schedule(dag, all, nis, G.x_slots[6]);
    }
  }
  bool class_mode_ = false;

  // ---- grad!: which sparse-gradient slots can be written without atomics ------------------------------
  // A slot whose index is  base ± (k+1)  over an unrestricted support index walks a contiguous block of
  // g exactly once; if no other objective slot touches that block, every entry has a single writer:
  // the kernel stores it directly and grad! does not have to zero the block first.  Everything else
  // (finite / shared variables, restricted or product iterators) keeps the warp-reduce + atomic path.
  std::vector<std::pair<int64_t, int64_t>> grad_zero_ranges; // 0-based [start, start+len) that grad! must zero
  void analyse_grad() {
    struct Iv { int64_t lo, hi; int gen, slot; bool simple; int coef; };
    std::vector<Iv> ivs;
    bool complex_any = false;
    for (size_t gi = 0; gi < objs.size(); ++gi) {
      Generator &g = objs[gi];
      g.grad_direct.assign(g.c.jac_slot.size(), 0);
      const Iterator &it = itrs[g.itr];
      for (size_t c = 0; c < g.c.jac_slot.size(); ++c) {
        const IndexExpr &e = g.c.uidx[g.c.jac_slot[c]];
        if (e.terms.empty()) { ivs.push_back(Iv{e.base, e.base, (int)gi, (int)c, false, 0}); continue; }
        bool ok = e.terms.size() == 1 && (e.terms[0].second == 1 || e.terms[0].second == -1);
        if (ok) {
          const ColRef &r = it.int_cols[g.c.int_cols[e.terms[0].first]];
          ok = columns[r.col].iota && r.div == 1 && r.mod == g.K && g.K > 0;
        }
        if (!ok) { complex_any = true; continue; }
        int64_t a = e.base + e.terms[0].second * 1, b = e.base + e.terms[0].second * g.K;
        ivs.push_back(Iv{std::min(a, b), std::max(a, b), (int)gi, (int)c, true, (int)e.terms[0].second});
      }
    }
    grad_zero_ranges.clear();
    if (!complex_any) {
      std::vector<size_t> order(ivs.size());
      for (size_t i = 0; i < order.size(); ++i) order[i] = i;
      std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return ivs[a].lo < ivs[b].lo; });
      std::vector<uint8_t> clash(ivs.size(), 0);
      int64_t reach = INT64_MIN; size_t reach_i = 0;
      for (size_t q = 0; q < order.size(); ++q) {
        size_t i = order[q];
        if (q > 0 && ivs[i].lo <= reach) { clash[i] = 1; clash[reach_i] = 1; }
        if (ivs[i].hi > reach) { reach = ivs[i].hi; reach_i = i; }
      }
      for (size_t i = 0; i < ivs.size(); ++i)
        if (ivs[i].simple && !clash[i] && ivs[i].lo >= 1 && ivs[i].hi <= nvar) objs[ivs[i].gen].grad_direct[ivs[i].slot] = 1;
    }
    // zero ranges: the complement (in [0, nvar)) of the blocks this RANK stores directly
    std::vector<std::pair<int64_t, int64_t>> direct;
    for (size_t i = 0; i < ivs.size() && !complex_any; ++i) {
      const Generator &g = objs[ivs[i].gen];
      if (!g.grad_direct[ivs[i].slot] || g.k1 <= g.k0) continue;
      const IndexExpr &e = g.c.uidx[g.c.jac_slot[ivs[i].slot]];
      int64_t a = e.base + ivs[i].coef * (g.k0 + 1), b = e.base + ivs[i].coef * g.k1; // 1-based, inclusive
      direct.push_back({std::min(a, b) - 1, std::max(a, b)});                           // 0-based [lo, hi)
    }
    std::sort(direct.begin(), direct.end());
    int64_t pos = 0;
    for (auto &d : direct) {
      if (d.first > pos) grad_zero_ranges.push_back({pos, d.first - pos});
      pos = std::max(pos, d.second);
    }
    if (pos < nvar) grad_zero_ranges.push_back({pos, nvar - pos});
    if (world > 1) {
      // sharded: a rank contributes nothing outside the parts of x it reads (its own supports, shared variables, halos) —
      // zero-filling the other ranks' 1 - 1/world of g on every call made grad! the one callback that did not scale
      // (8 GPUs, config 3: 0.068 ms of which 0.05 ms were the memset).  Entries outside iexa_x_ranges are left UNTOUCHED.
      std::vector<std::pair<int64_t, int64_t>> rr = x_read_ranges(), cut; // 1-based inclusive, sorted, disjoint
      size_t j = 0;
      for (auto &z : grad_zero_ranges) {
        const int64_t zlo = z.first, zhi = z.first + z.second; // 0-based [zlo, zhi)
        while (j < rr.size() && rr[j].second <= zlo) ++j;      // rr[j] as 0-based [first-1, second)
        for (size_t q = j; q < rr.size() && rr[q].first - 1 < zhi; ++q) {
          const int64_t lo = std::max(zlo, rr[q].first - 1), hi = std::min(zhi, rr[q].second);
          if (lo < hi) cut.push_back({lo, hi - lo});
        }
      }
      grad_zero_ranges.swap(cut);
    }
  }


  // ---- index-expression queries (sharding, single-writer analysis, bounds checks) ------------------------
  // conservative cover [lo, hi] (1-based, inclusive) of  base + sum coef*col(k)  over supports k in [k0, k1)
  void index_range(const Iterator &it, const std::vector<int32_t> &int_cols, const IndexExpr &e, int64_t k0, int64_t k1,
                   int64_t &lo, int64_t &hi) const {
    lo = hi = e.base;
    for (auto &t : e.terms) {
      const ColRef &r = it.int_cols[int_cols[t.first]];
      const HostColumn &c = columns[r.col];
      // positions j = (k / div) % mod visited by k in [k0, k1)
      int64_t q0 = k0 / r.div, q1 = (k1 - 1) / r.div, j0 = 0, j1 = r.mod - 1;
      if (q1 - q0 + 1 < r.mod && q0 % r.mod <= q1 % r.mod) { j0 = q0 % r.mod; j1 = q1 % r.mod; }
      int64_t vmin, vmax;
      if (c.affine) {
        const int64_t a0 = c.ab * (j0 / c.ac), a1 = c.ab * (j1 / c.ac), dm = c.ad * (c.ac - 1);
        vmin = c.aa + std::min(a0, a1) + std::min<int64_t>(0, dm);
        vmax = c.aa + std::max(a0, a1) + std::max<int64_t>(0, dm);
      } else if (j0 == 0 && j1 == (int64_t)c.ivals.size() - 1) {
        vmin = c.vmin; vmax = c.vmax;
      } else {
        vmin = vmax = c.ivals[j0];
        for (int64_t j = j0; j <= j1; ++j) { vmin = std::min<int64_t>(vmin, c.ivals[j]); vmax = std::max<int64_t>(vmax, c.ivals[j]); }
      }
      lo += t.second >= 0 ? t.second * vmin : t.second * vmax;
      hi += t.second >= 0 ? t.second * vmax : t.second * vmin;
    }
  }
  // index(k) == c0 + s*k for EVERY k in [0, K) with s = +1 or -1?  True for an unrestricted support index
  // (base + k + 1), for shifted ranges (2..T), and for the column-major index of a variable over a product iterator
  // whose factors are all unrestricted (i_t + T*(i_xi - 1): the mixed-radix digits of k recombine to k itself).
  bool linear_in_k(const Iterator &it, const std::vector<int32_t> &int_cols, int64_t K, const IndexExpr &e,
                   int64_t &c0, int64_t &s) const {
    if (e.terms.empty() || K <= 0) return false;
    std::map<std::pair<int64_t, int64_t>, int64_t> w; // digit (div, mod) of k -> weight
    c0 = e.base;
    for (auto &t : e.terms) {
      const ColRef &r = it.int_cols[int_cols[t.first]];
      const HostColumn &c = columns[r.col];
      if (!c.affine || c.ac != 1) return false;
      c0 += t.second * c.aa;
      if (r.mod > 1) w[{r.div, r.mod}] += t.second * c.ab;
    }
    int64_t expect_div = 1, last = 0;
    s = 0;
    for (auto &kv : w) { // ascending div
      if (kv.first.first != expect_div) return false;
      if (s == 0) { s = kv.second; if (s != 1 && s != -1) return false; }
      if (kv.second != s * kv.first.first) return false;
      expect_div = kv.first.first * kv.first.second;
      last = expect_div;
    }
    return s != 0 && last >= K;
  }

  // the parts of x this rank's callbacks READ (1-based inclusive, merged, sorted): its own supports of every variable block,
  // the replicated finite / shared variables and the halo of shifted references at the shard boundaries (a cover)
  std::vector<std::pair<int64_t, int64_t>> x_read_ranges() const {
    std::vector<std::pair<int64_t, int64_t>> iv;
    auto scan = [&](const Generator &g) {
      if (g.k1 <= g.k0) return;
      const Iterator &it = itrs[g.itr];
      std::vector<uint8_t> used(g.c.uidx.size(), 0);
      for (size_t s = 0; s < used.size(); ++s)
        used[s] = (s < g.c.x_slots_val.size() && g.c.x_slots_val[s]) || (s < g.c.x_slots_d1.size() && g.c.x_slots_d1[s]) ||
                  (s < g.c.x_slots_d2.size() && g.c.x_slots_d2[s]);
      for (int32_t s : g.c.jac_slot) used[s] = 1;
      for (auto &pr : g.c.hess_slot) { used[pr.first] = 1; used[pr.second] = 1; }
      for (size_t s = 0; s < used.size(); ++s) {
        if (!used[s]) continue;
        int64_t lo, hi;
        index_range(it, g.c.int_cols, g.c.uidx[s], g.k0, g.k1, lo, hi);
        lo = std::max<int64_t>(lo, 1); hi = std::min<int64_t>(hi, nvar);
        if (lo <= hi) iv.emplace_back(lo, hi);
      }
    };
    for (auto &g : objs) scan(g);
    for (auto &g : cons) scan(g);
    std::sort(iv.begin(), iv.end());
    std::vector<std::pair<int64_t, int64_t>> m;
    for (auto &r : iv) {
      if (!m.empty() && r.first <= m.back().second + 1) m.back().second = std::max(m.back().second, r.second);
      else m.push_back(r);
    }
    return m;
  }

  // ---- what a rank keeps on ITS device when world > 1 (shard the inputs, not just the outputs) ----------
  // positions [lo, hi) of column c that this rank's generators visit: j = (k / div) % mod over k in [k0, k1) for every
  // reference to the column (a hull; a wrapping or longer-than-mod walk keeps the whole column).  lo == hi: not read at all.
  // Parameter functions are evaluated over the k-range whose theta entries this rank reads (shard_pfuncs).
  std::vector<std::pair<int64_t, int64_t>> column_read_ranges() const {
    std::vector<std::pair<int64_t, int64_t>> rr(columns.size(), {INT64_MAX, 0});
    auto visit = [&](const ColRef &r, int64_t k0, int64_t k1) {
      if (k1 <= k0) return;
      int64_t q0 = k0 / r.div, q1 = (k1 - 1) / r.div, j0 = 0, j1 = r.mod - 1;
      if (q1 - q0 + 1 < r.mod && q0 % r.mod <= q1 % r.mod) { j0 = q0 % r.mod; j1 = q1 % r.mod; }
      rr[r.col].first = std::min(rr[r.col].first, j0);
      rr[r.col].second = std::max(rr[r.col].second, j1 + 1);
    };
    auto scan = [&](const Generator &g) {
      const Iterator &it = itrs[g.itr];
      for (int32_t s : g.c.int_cols) visit(it.int_cols[s], g.k0, g.k1);
      for (int32_t s : g.c.fp_cols) visit(it.fp_cols[s], g.k0, g.k1);
    };
    for (auto &g : objs) scan(g);
    for (auto &g : cons) scan(g);
    for (auto &g : pfuncs) scan(g);
    for (size_t c = 0; c < columns.size(); ++c) {
      if (rr[c].first == INT64_MAX) rr[c] = {0, 0};
      const HostColumn &hc = columns[c];
      if (!hc.is_int && hc.gen_kind == 3 && hc.gen_src >= 0 && rr[c].second > rr[c].first) { // TRAPEZOID reads src[j-1 .. j+1]
        auto &sr = rr[hc.gen_src];
        const int64_t lo = std::max<int64_t>(rr[c].first - 1, 0), hi = std::min<int64_t>(rr[c].second + 1, columns[hc.gen_src].K);
        if (sr.second <= sr.first) sr = {lo, hi};
        else { sr.first = std::min(sr.first, lo); sr.second = std::max(sr.second, hi); }
      }
    }
    return rr;
  }
  // theta entries [lo, hi) (0-based) that the programs of generator g read over its supports [k0, k1) (a cover)
  void theta_reads_of(const Generator &g, std::vector<std::pair<int64_t, int64_t>> &iv) const {
    if (g.k1 <= g.k0) return;
    const Iterator &it = itrs[g.itr];
    std::vector<uint8_t> used(g.c.uidx.size(), 0);
    const Program *pr[] = {&g.c.val, &g.c.d1, &g.c.d2, &g.c.jv, &g.c.jtv, &g.c.hv};
    for (const Program *q : pr) for (const Instr &I : q->code) if (I.op == D_LOADP && I.a >= 0 && (size_t)I.a < used.size()) used[I.a] = 1;
    for (size_t s2 = 0; s2 < used.size(); ++s2) {
      if (!used[s2]) continue;
      int64_t lo, hi;
      index_range(it, g.c.int_cols, g.c.uidx[s2], g.k0, g.k1, lo, hi);
      lo = std::max<int64_t>(lo, 1); hi = std::min<int64_t>(hi, npar);
      if (lo <= hi) iv.emplace_back(lo - 1, hi);
    }
  }
  // world > 1: a parameter function (a theta block evaluated on the device at finalize) runs over the k-range whose entries this
  // rank's generators — or a later parameter function, through PAR leaves — read; the rest of the block belongs to other ranks.
  // Walked last to first, so that what a restricted function reads is known when the earlier ones are restricted.
  void shard_pfuncs() {
    for (auto &g : pfuncs) { g.k0 = 0; g.k1 = g.K; g.l0 = g.o0; }
    if (world <= 1 || pfuncs.empty() || getenv("IEXA_NO_PFUNC_SHARDING")) return;
    std::vector<std::pair<int64_t, int64_t>> reads;
    for (auto &g : objs) theta_reads_of(g, reads);
    for (auto &g : cons) theta_reads_of(g, reads);
    for (size_t pi = pfuncs.size(); pi-- > 0;) {
      Generator &g = pfuncs[pi];
      int64_t lo = INT64_MAX, hi = 0;
      for (auto &r : reads) {
        const int64_t a = std::max(r.first, g.o0), b = std::min(r.second, g.o0 + g.K);
        if (a < b) { lo = std::min(lo, a); hi = std::max(hi, b); }
      }
      if (lo == INT64_MAX) { g.k0 = g.k1 = 0; }
      else { g.k0 = lo - g.o0; g.k1 = hi - g.o0; }
      g.l0 = g.o0 + g.k0;
      theta_reads_of(g, reads);
    }
  }
  // the parts of theta resident on this rank (0-based [lo, hi), merged, sorted; a cover): what its generators read and what its
  // share of every parameter function reads and writes
  std::vector<std::pair<int64_t, int64_t>> theta_read_ranges() const {
    std::vector<std::pair<int64_t, int64_t>> iv, m;
    if (npar <= 0) return m;
    if (world <= 1) { m.emplace_back(0, npar); return m; }
    for (auto &g : objs) theta_reads_of(g, iv);
    for (auto &g : cons) theta_reads_of(g, iv);
    for (auto &g : pfuncs) {
      theta_reads_of(g, iv);
      if (g.k1 > g.k0) iv.emplace_back(g.o0 + g.k0, g.o0 + g.k1);
    }
    std::sort(iv.begin(), iv.end());
    for (auto &r : iv) {
      if (!m.empty() && r.first <= m.back().second) m.back().second = std::max(m.back().second, r.second);
      else m.push_back(r);
    }
    return m;
  }

  // IEXA_SLOT_ORDER_JAC_ROW_SORTED.  perm[p] = policy-order slot that comes p-th when the first-order slots of a row are sorted by
  // column index — provided that order is STATIC (the same for every support k in [0, K)) and strict (no two slots share a
  // column at any k).  Pairs of slots with identical term structure differ by a constant; every other pair is checked over
  // all supports.  Then rows are contiguous and column-sorted in the COO array: it IS the CSR value array, with
  // rowptr[o0 + k] = o1 + o1step*k — the COO->CSR pass of the Jacobian disappears (iexa_jac_rowptr).
  bool jac_row_sort(const Iterator &it, const SlotCtx &ctx, const std::vector<int32_t> &jac_slot, std::vector<int32_t> &perm) const {
    const size_t n = jac_slot.size();
    perm.resize(n);
    for (size_t i = 0; i < n; ++i) perm[i] = (int32_t)i;
    if (n <= 1) return true;
    if (it.K <= 0) return true;
    auto col_at = [&](int32_t slot, int64_t k) {
      const ColRef &r = it.int_cols[ctx.int_cols[slot]];
      return columns[r.col].ival((k / r.div) % r.mod);
    };
    auto idx_at = [&](const IndexExpr &e, int64_t k) {
      int64_t v = e.base;
      for (auto &t : e.terms) v += t.second * col_at(t.first, k);
      return v;
    };
    std::vector<int64_t> key0(n);
    for (size_t i = 0; i < n; ++i) key0[i] = idx_at(ctx.uidx[jac_slot[i]], 0);
    std::stable_sort(perm.begin(), perm.end(), [&](int32_t a, int32_t b) { return key0[a] < key0[b]; });
    for (size_t p = 0; p + 1 < n; ++p) {
      const IndexExpr &ea = ctx.uidx[jac_slot[perm[p]]], &eb = ctx.uidx[jac_slot[perm[p + 1]]];
      if (key0[perm[p]] >= key0[perm[p + 1]]) return false;      // duplicate column at k = 0
      if (ea.terms == eb.terms) continue;                         // constant difference: static
      // The answer depends on the iterator, on the two term lists (in the iterator's own column numbering) and on the
      // difference of the bases only: the nine state rows of a collocation scheme ask the same question nine times
      std::string key;
      auto put = [&](int64_t v) { key.append(reinterpret_cast<const char *>(&v), sizeof v); };
      put((int64_t)(&it - itrs.data())); put(eb.base - ea.base);
      for (const IndexExpr *e : {&ea, &eb}) { put((int64_t)e->terms.size()); for (auto &t : e->terms) { put(ctx.int_cols[t.first]); put(t.second); } }
      auto hit = row_sort_memo_.find(key);
      bool ok;
      if (hit != row_sort_memo_.end()) ok = hit->second;
      else {
        ok = true;
        for (int64_t k = 1; k < it.K && ok; ++k)                  // neighbours in the sorted order stay strictly ordered
          ok = idx_at(ea, k) < idx_at(eb, k);
        row_sort_memo_[key] = ok;
      }
      if (!ok) return false;
    }
    return true;
  }
  mutable std::map<std::string, bool> row_sort_memo_;
  // every constraint generator's rows are column-sorted (policy IEXA_SLOT_ORDER_JAC_ROW_SORTED and a static order exists)
  bool jac_is_csr() const {
    for (const Generator &g : cons) if (g.c.o1step > 1 && !g.c.jac_row_sorted) return false;
    return opt_slot_order == 2;
  }

  // 0-based row pointers of this rank's rows: the rows of a generator are o1step slots apart (structure, built once per model)
  std::vector<int64_t> jac_rowptr() const {
    std::vector<int64_t> rp((size_t)loc_ncon + 1);
    for (const Generator &g : cons)
      for (int64_t k = g.k0; k < g.k1; ++k) rp[g.l0 + (k - g.k0)] = g.l1 + (int64_t)g.c.o1step * (k - g.k0);
    rp[(size_t)loc_ncon] = loc_nnzj;
    return rp;
  }

  // every x / theta index a tape can produce must stay inside [1, nvar] / [1, npar]: the kernels address
  // x[idx-1] / theta[idx-1] / out[idx-1] unchecked (a bad tape must be IEXA_ERR_INVALID, not an Xid)
  void check_index_bounds() const {
    auto check = [&](const Generator &g) {
      if (g.K <= 0) return;
      const Iterator &it = itrs[g.itr];
      std::vector<int8_t> kind(g.c.uidx.size(), 0); // bit 0: VAR leaf, bit 1: PAR leaf
      for (const iexa_node &n : g.c.tape) {
        if (n.op == IEXA_OP_VAR) kind[g.c.idx_map[n.a]] |= 1;
        else if (n.op == IEXA_OP_PAR) kind[g.c.idx_map[n.a]] |= 2;
      }
      for (size_t u = 0; u < kind.size(); ++u) {
        for (int bit = 1; bit <= 2; bit <<= 1) {
          if (!(kind[u] & bit)) continue;
          const int64_t n = bit == 1 ? nvar : npar;
          int64_t lo, hi;
          index_range(it, g.c.int_cols, g.c.uidx[u], 0, g.K, lo, hi);
          if (lo >= 1 && hi <= n) continue;
          // the cover is conservative for correlated columns: decide exactly before rejecting
          for (int64_t k = 0; k < g.K; ++k) {
            const int64_t v = index_value(g, (int32_t)u, k);
            if (v < 1 || v > n)
              throw std::invalid_argument(std::string(bit == 1 ? "variable" : "parameter") + " index " + std::to_string(v) +
                                          " out of range [1, " + std::to_string(n) + "] at support " + std::to_string(k + 1) +
                                          " of " + (g.is_obj ? "an objective" : "a constraint") + " generator");
          }
        }
      }
    };
    for (auto &g : objs) check(g);
    for (auto &g : cons) check(g);
  }

  // ---- jtprod! / hprod!: two-phase scatter ----------------------------------------------------------------
  // The scatter products add into a dense nvar-vector.  An output of a group whose index is  c0 ± k  walks a contiguous
  // block exactly once (single writer INSIDE the group).  Groups are visited longest first; a group joins PHASE 0 when
  // its single-writer blocks touch nothing an already accepted phase-0 group writes (directly or atomically) and its
  // other outputs stay clear of every accepted block: its blocks are then plain stores that need no zero-fill.
  // Everything else runs in PHASE 1 — a second launch, ordered after the first — with atomics (warp-reduced for shared
  // variables).  Quadrotor jtprod!: the fused ODE rows (K = T) store all 22 variable blocks directly, the collocation /
  // restriction / initial-condition rows add on top in phase 1, nothing is zero-filled.
  std::vector<std::pair<int64_t, int64_t>> scat_zero_ranges[2]; // [jtprod, hprod]: 0-based (start, length) zeroed first
  void analyse_scatter() {
    const bool no_direct = getenv("IEXA_NO_SCATTER_DIRECT") != nullptr;
    typedef std::pair<int64_t, int64_t> Iv; // [lo, hi] 1-based inclusive
    auto overlaps = [](const std::vector<Iv> &v, const Iv &a) {
      for (const Iv &b : v) if (a.first <= b.second && b.first <= a.second) return true;
      return false;
    };
    for (int w = 0; w < 2; ++w) {
      const int prog = 4 + w;
      std::vector<int> order;
      for (size_t gi = 0; gi < groups.size(); ++gi) {
        Group &G = groups[gi];
        G.scat_phase[w] = 1;
        G.scat_rider_of[w] = -1;
        G.scat_direct[w].assign(G.outmap[prog].size(), 0);
        if (G.prog[prog].nout > 0 && G.k1 > G.k0 && !G.is_class && !no_direct) order.push_back((int)gi);
      }
      std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return groups[a].k1 - groups[a].k0 > groups[b].k1 - groups[b].k0; });
      std::vector<Iv> direct, atomic;
      for (int gi : order) {
        Group &G = groups[gi];
        const Iterator &it = itrs[G.itr];
        std::vector<Iv> D, A;
        std::vector<uint8_t> isd(G.outmap[prog].size(), 0);
        for (size_t j = 0; j < G.outmap[prog].size(); ++j) {
          const IndexExpr &e = G.ctx.uidx[G.outmap[prog][j].second];
          int64_t c0, sgn, lo, hi;
          if (linear_in_k(it, G.ctx.int_cols, G.K, e, c0, sgn)) {
            const int64_t a = c0 + sgn * G.k0, b = c0 + sgn * (G.k1 - 1);
            D.push_back({std::min(a, b), std::max(a, b)});
            isd[j] = 1;
          } else {
            index_range(it, G.ctx.int_cols, e, G.k0, G.k1, lo, hi);
            A.push_back({lo, hi});
          }
        }
        if (D.empty()) continue;
        bool ok = true;
        for (size_t i = 0; i < D.size() && ok; ++i) {
          if (D[i].first < 1 || D[i].second > nvar) ok = false;
          for (size_t j = i + 1; j < D.size() && ok; ++j) if (D[i].first <= D[j].second && D[j].first <= D[i].second) ok = false;
          if (ok && (overlaps(direct, D[i]) || overlaps(atomic, D[i]) || overlaps(A, D[i]))) ok = false;
        }
        for (size_t i = 0; i < A.size() && ok; ++i) if (overlaps(direct, A[i])) ok = false;
        if (!ok) continue;
        G.scat_phase[w] = 0;
        G.scat_direct[w] = isd;
        direct.insert(direct.end(), D.begin(), D.end());
        atomic.insert(atomic.end(), A.begin(), A.end());
      }
      // riders (see Group::scat_rider_of)
      if (!getenv("IEXA_NO_RIDERS")) {
        struct St { int64_t c0, s; Iv iv; };
        std::map<int, std::vector<St>> stored; // primary -> what its unit stores, thread k <-> entry c0 + s*k
        auto stores_of = [&](int gi) -> std::vector<St> & {
          auto it = stored.find(gi);
          if (it != stored.end()) return it->second;
          std::vector<St> v;
          const Group &G = groups[gi];
          const Iterator &itr = itrs[G.itr];
          for (size_t j = 0; j < G.outmap[prog].size(); ++j) {
            if (!G.scat_direct[w][j]) continue;
            int64_t c0, sg;
            if (!linear_in_k(itr, G.ctx.int_cols, G.K, G.ctx.uidx[G.outmap[prog][j].second], c0, sg)) continue;
            const int64_t a = c0 + sg * G.k0, b = c0 + sg * (G.k1 - 1);
            v.push_back(St{c0, sg, {std::min(a, b), std::max(a, b)}});
          }
          return stored.emplace(gi, std::move(v)).first->second;
        };
        for (size_t hi = 0; hi < groups.size(); ++hi) {
          Group &H = groups[hi];
          if (H.is_class || H.scat_phase[w] == 0 || H.prog[prog].nout == 0 || H.k1 <= H.k0) continue;
          for (int gi : order) {
            const Group &G = groups[gi];
            if (G.scat_phase[w] != 0 || G.scat_rider_of[w] >= 0 || G.K != H.K || G.k0 != H.k0 || G.k1 != H.k1 || (size_t)gi == hi) continue; // same support count: thread k <-> support k of both
            std::vector<St> &unit = stores_of(gi);
            const Iterator &itr = itrs[H.itr];
            std::vector<uint8_t> mode(H.outmap[prog].size(), 0);
            std::vector<St> fresh;
            std::vector<Iv> Aat;
            bool ok = true, any = false;
            for (size_t j = 0; j < H.outmap[prog].size() && ok; ++j) {
              const IndexExpr &e = H.ctx.uidx[H.outmap[prog][j].second];
              int64_t c0, sg, lo, hi2;
              if (linear_in_k(itr, H.ctx.int_cols, H.K, e, c0, sg)) {
                const int64_t a = c0 + sg * H.k0, b = c0 + sg * (H.k1 - 1);
                const Iv iv{std::min(a, b), std::max(a, b)};
                bool same = false;
                for (const St &u : unit) same = same || (u.c0 == c0 && u.s == sg);
                for (const St &u : fresh) same = same || (u.c0 == c0 && u.s == sg);
                if (same) { mode[j] = 2; any = true; }
                else if (iv.first >= 1 && iv.second <= nvar && !overlaps(direct, iv) && !overlaps(atomic, iv)) {
                  bool clash = false;
                  for (const St &u : fresh) clash = clash || (iv.first <= u.iv.second && u.iv.first <= iv.second);
                  for (const Iv &a2 : Aat) clash = clash || (iv.first <= a2.second && a2.first <= iv.second);
                  if (clash) ok = false; else { mode[j] = 1; fresh.push_back(St{c0, sg, iv}); any = true; }
                } else ok = false;
              } else {
                index_range(itr, H.ctx.int_cols, e, H.k0, H.k1, lo, hi2);
                const Iv iv{lo, hi2};
                bool clash = overlaps(direct, iv);
                for (const St &u : fresh) clash = clash || (iv.first <= u.iv.second && u.iv.first <= iv.second);
                if (clash) ok = false; else Aat.push_back(iv);
              }
            }
            if (!ok || !any) continue;
            H.scat_phase[w] = 0;
            H.scat_rider_of[w] = gi;
            H.scat_direct[w] = mode;
            for (const St &u : fresh) { unit.push_back(u); direct.push_back(u.iv); }
            atomic.insert(atomic.end(), Aat.begin(), Aat.end());
            break;
          }
        }
      }
      std::sort(direct.begin(), direct.end());
      scat_zero_ranges[w].clear();
      int64_t pos = 0; // 0-based
      for (const Iv &d : direct) {
        if (d.first - 1 > pos) scat_zero_ranges[w].push_back({pos, d.first - 1 - pos});
        pos = std::max(pos, d.second);
      }
      if (pos < nvar) scat_zero_ranges[w].push_back({pos, nvar - pos});
    }
  }

  const Generator &member(const Group &G, int mpos) const { return (G.is_obj ? objs : cons)[G.members[mpos]]; }
};

} // namespace iexa

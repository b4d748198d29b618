/*
 * oracle.c — TEST INFRASTRUCTURE ONLY.  CPU restatement of the evaluation algorithm that the
 * reference reaches through ExaModels.jl (pinned "0.11.2", /root/reference/Project.toml:23).
 *
 * PARITY UNPINNED: ExaModels.jl is a third-party Julia package that is NOT vendored under
 * /root/reference and cannot be installed or run here (no Julia, no network).  None of the
 * reference's own tests call cons!/jac_coord!/hess_coord! directly; the only
 * pinned numbers are solve-level objective values and problem dimensions
 * (/root/reference/test/madnlp.jl:18,42, test/ipopt.jl:183-186, test/solve.jl:146,154,187,206),
 * which tests/test_golden_solves.py reproduces THROUGH this oracle.  The algorithm below restates
 * ExaModels' published design (src/graph.jl, simdfunction.jl, gradient.jl, jacobian.jl,
 * hessian.jl, nlp.jl as recalled in SURVEY.md Appendix A): per support point, build the tree of
 * local partials, then run the recursive reverse passes
 *     drpass / grpass / jrpass   (first order,  Appendix A.3)
 *     hrpass0 / hrpass / hdrpass (second order, Appendix A.4)
 * writing COO entries at   o + ostep*(k-1) + comp(cnt).
 * It deliberately shares NO code with the product (infiniteexamodels.jl_b200/csrc): the product
 * differentiates symbolically once per generator into a DAG; this file recurses numerically over
 * the tree at every support like ExaModels does.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.
 *
 * Call sites in the reference whose semantics are followed:
 *   add_con  src/transform.jl:458,559,597     add_obj  src/transform.jl:614,700,741
 *   operator set  src/operators.jl:2-46       Null  src/transform.jl:393
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* operator codes: numerically identical to include/iexa.h so one Python lowering feeds both */
enum {
  OP_CONST = 0, OP_FIELD = 1, OP_VAR = 2, OP_PAR = 3,
  OP_ADD = 10, OP_SUB, OP_MUL, OP_DIV, OP_POW,
  OP_NEG = 20, OP_POS, OP_INV, OP_SQRT, OP_CBRT, OP_ABS, OP_ABS2, OP_EXP, OP_EXP2, OP_LOG,
  OP_LOG2, OP_LOG10, OP_LOG1P, OP_SIN, OP_COS, OP_TAN, OP_ASIN, OP_ACOS, OP_CSC, OP_SEC,
  OP_COT, OP_ATAN, OP_ACOT, OP_SIND, OP_COSD, OP_TAND, OP_CSCD, OP_SECD, OP_COTD, OP_ATAND,
  OP_ACOTD, OP_SINH, OP_COSH, OP_TANH, OP_CSCH, OP_SECH, OP_COTH, OP_ATANH, OP_ACOTH
};

typedef struct { int32_t op, a, b, pad; double c; } onode;
typedef struct { int64_t base; int32_t nterms; int32_t col[4]; int32_t pad; int64_t coef[4]; } oindex;

/* node kinds after the static "contains a Var" analysis (constant operands make unary nodes,
 * Appendix A.3: "Binary ops with a constant operand are stored as unary nodes") */
enum { KD_CONST = 0, KD_VAR = 1, KD_UN = 2, KD_BIN = 3 };

typedef struct {
  int n; onode *nd; int *kind; int *c1; int *c2;
  int n_idx; oindex *idx; int *canon; /* canon[i] = first index id with the same affine form */
  int64_t K; int n_int; int64_t **ic; int n_fp; double **fc;
  int is_obj; double lcon, ucon;
  int64_t o0, o1, o2;
  int o1step, o2step, nocc1, nocc2;
  int *comp1, *comp2;       /* occurrence -> compressed slot (0-based) */
  int *slot1_idx;           /* slot -> canonical index id */
  int *slot2_i, *slot2_j;
  int borrowed;             /* iterator columns belong to the caller (orc_add_gen_borrow) */
  int order;                /* slot-order policy: 0 children left to right, 1 right to left (Appendix A.2 is a hypothesis) */
} ogen;

typedef struct {
  int ngen, cap; ogen *g;
  int64_t nvar, ncon, nnzj, nnzh, npar;
  double *theta;
  int order;                /* policy given to generators added from now on */
  int row_sorted;           /* policy 2: constraint generators' first-order slots column-sorted when a static order exists */
  int all_sorted;           /* every constraint generator added under policy 2 had such an order */
} omodel;

/* ------------------------------------------------------------------------------------------ */
static int idx_same(const oindex *a, const oindex *b) {
  /* identical index EXPRESSIONS (Appendix A.2), compared as canonical affine forms */
  if (a->base != b->base) return 0;
  int64_t ca[64], cb[64];
  memset(ca, 0, sizeof ca); memset(cb, 0, sizeof cb);
  for (int t = 0; t < a->nterms; ++t) ca[a->col[t] & 63] += a->coef[t];
  for (int t = 0; t < b->nterms; ++t) cb[b->col[t] & 63] += b->coef[t];
  return memcmp(ca, cb, sizeof ca) == 0;
}

static int64_t idx_val(const ogen *g, int id, int64_t k) {
  const oindex *e = &g->idx[id];
  int64_t v = e->base;
  for (int t = 0; t < e->nterms; ++t) v += e->coef[t] * g->ic[e->col[t]][k];
  return v;
}

/* thread count of the OpenMP loops over supports (1 = the sequential loops of ExaModels' CPU backend) */
void orc_set_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n > 0 ? n : 1);
#else
  (void)n;
#endif
}
int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

omodel *orc_create(void) { return (omodel *)calloc(1, sizeof(omodel)); }
/* slot-order policy of the generators added AFTER this call: 0 = inner1 then inner2, 1 = inner2 then inner1 */
void orc_set_slot_order(omodel *m, int order) { m->order = order == 1; m->row_sorted = order == 2; }
int orc_jac_is_csr(const omodel *m) { return m->row_sorted && m->all_sorted == 0; }
/* first / second child of a binary node in the policy's visiting order */
#define KID1(g, t) ((g)->order ? (g)->c2[t] : (g)->c1[t])
#define KID2(g, t) ((g)->order ? (g)->c1[t] : (g)->c2[t])

void orc_free(omodel *m) {
  if (!m) return;
  for (int i = 0; i < m->ngen; ++i) {
    ogen *g = &m->g[i];
    free(g->nd); free(g->kind); free(g->c1); free(g->c2); free(g->idx); free(g->canon);
    for (int j = 0; j < g->n_int && !g->borrowed; ++j) free(g->ic[j]);
    for (int j = 0; j < g->n_fp && !g->borrowed; ++j) free(g->fc[j]);
    free(g->ic); free(g->fc); free(g->comp1); free(g->comp2);
    free(g->slot1_idx); free(g->slot2_i); free(g->slot2_j);
  }
  free(m->g); free(m->theta); free(m);
}

void orc_set_dims(omodel *m, int64_t nvar, int64_t npar, const double *theta) {
  m->nvar = nvar; m->npar = npar;
  free(m->theta);
  m->theta = (double *)malloc(sizeof(double) * (size_t)(npar > 0 ? npar : 1));
  if (npar > 0) memcpy(m->theta, theta, sizeof(double) * (size_t)npar);
}
void orc_set_theta(omodel *m, int64_t off, int64_t n, const double *v) {
  memcpy(m->theta + off, v, sizeof(double) * (size_t)n);
}

/* ---- symbolic passes (SIMDFunction construction, Appendix A.2) ---------------------------- */
typedef struct { int *a; int *b; int n, cap; } plist;
static void pl_push(plist *l, int a, int b) {
  if (l->n == l->cap) {
    l->cap = l->cap ? 2 * l->cap : 64;
    l->a = (int *)realloc(l->a, sizeof(int) * (size_t)l->cap);
    l->b = (int *)realloc(l->b, sizeof(int) * (size_t)l->cap);
  }
  l->a[l->n] = a; l->b[l->n] = b; l->n++;
}

static void sym_first(const ogen *g, int t, plist *out) {
  switch (g->kind[t]) {
    case KD_VAR: pl_push(out, g->canon[g->nd[t].a], 0); break;
    case KD_UN: sym_first(g, g->c1[t], out); break;
    case KD_BIN: sym_first(g, KID1(g, t), out); sym_first(g, KID2(g, t), out); break;
    default: break;
  }
}
/* children of a node for the cross pass: a Var leaf stands for itself */
static int kids(const ogen *g, int t, int *c) {
  switch (g->kind[t]) {
    case KD_VAR: c[0] = t; return 1;
    case KD_UN: c[0] = g->c1[t]; return 1;
    case KD_BIN: c[0] = KID1(g, t); c[1] = KID2(g, t); return 2;
    default: return 0;
  }
}
static void sym_hdr(const ogen *g, int t1, int t2, plist *out) {
  if (g->kind[t1] == KD_VAR && g->kind[t2] == KD_VAR) { pl_push(out, g->canon[g->nd[t1].a], g->canon[g->nd[t2].a]); return; }
  int a[2], b[2], na = kids(g, t1, a), nb = kids(g, t2, b);
  for (int i = 0; i < na; ++i) for (int j = 0; j < nb; ++j) sym_hdr(g, a[i], b[j], out);
}
static void sym_hr(const ogen *g, int t, plist *out) {
  switch (g->kind[t]) {
    case KD_VAR: { int u = g->canon[g->nd[t].a]; pl_push(out, u, u); break; }
    case KD_UN: sym_hr(g, g->c1[t], out); break;
    case KD_BIN: sym_hr(g, KID1(g, t), out); sym_hr(g, KID2(g, t), out); sym_hdr(g, KID1(g, t), KID2(g, t), out); break;
    default: break;
  }
}
/* is node t transparent for hrpass0?  (+, -, unary +/-, constant*subtree: Appendix A.4) */
static int passthrough(const ogen *g, int t) {
  int op = g->nd[t].op, kd = g->kind[t];
  if (kd == KD_BIN) return op == OP_ADD || op == OP_SUB;
  if (kd == KD_UN) return op == OP_ADD || op == OP_SUB || op == OP_MUL || op == OP_NEG || op == OP_POS;
  return 0;
}
static void sym_hr0(const ogen *g, int t, plist *out) {
  if (g->kind[t] == KD_VAR || g->kind[t] == KD_CONST) return;
  if (passthrough(g, t)) {
    if (g->kind[t] == KD_BIN) { sym_hr0(g, KID1(g, t), out); sym_hr0(g, KID2(g, t), out); }
    else sym_hr0(g, g->c1[t], out);
    return;
  }
  sym_hr(g, t, out);
}

static int compress(const plist *l, int **comp, int **sa, int **sb) {
  int ns = 0;
  *comp = (int *)malloc(sizeof(int) * (size_t)(l->n + 1));
  *sa = (int *)malloc(sizeof(int) * (size_t)(l->n + 1));
  *sb = (int *)malloc(sizeof(int) * (size_t)(l->n + 1));
  for (int i = 0; i < l->n; ++i) {
    int f = -1;
    for (int s = 0; s < ns; ++s) if ((*sa)[s] == l->a[i] && (*sb)[s] == l->b[i]) { f = s; break; }
    if (f < 0) { f = ns; (*sa)[ns] = l->a[i]; (*sb)[ns] = l->b[i]; ns++; }
    (*comp)[i] = f;
  }
  return ns;
}

static int add_gen(omodel *m, int is_obj, const onode *nodes, int n_nodes, const oindex *idx, int n_idx,
                   int64_t K, int n_int, const int64_t *const *icols, int n_fp, const double *const *fcols,
                   double lcon, double ucon, int borrow) {
  if (m->ngen == m->cap) { m->cap = m->cap ? 2 * m->cap : 16; m->g = (ogen *)realloc(m->g, sizeof(ogen) * (size_t)m->cap); }
  ogen *g = &m->g[m->ngen];
  memset(g, 0, sizeof *g);
  g->n = n_nodes; g->is_obj = is_obj; g->K = K; g->lcon = lcon; g->ucon = ucon; g->order = m->order;
  g->nd = (onode *)malloc(sizeof(onode) * (size_t)n_nodes); memcpy(g->nd, nodes, sizeof(onode) * (size_t)n_nodes);
  g->n_idx = n_idx;
  g->idx = (oindex *)malloc(sizeof(oindex) * (size_t)(n_idx + 1)); if (n_idx) memcpy(g->idx, idx, sizeof(oindex) * (size_t)n_idx);
  g->canon = (int *)malloc(sizeof(int) * (size_t)(n_idx + 1));
  for (int i = 0; i < n_idx; ++i) { g->canon[i] = i; for (int j = 0; j < i; ++j) if (idx_same(&idx[i], &idx[j])) { g->canon[i] = j; break; } }
  g->n_int = n_int; g->n_fp = n_fp;
  g->ic = (int64_t **)malloc(sizeof(void *) * (size_t)(n_int + 1));
  g->fc = (double **)malloc(sizeof(void *) * (size_t)(n_fp + 1));
  g->borrowed = borrow;
  for (int j = 0; j < n_int; ++j) {
    if (borrow) { g->ic[j] = (int64_t *)icols[j]; continue; }
    g->ic[j] = (int64_t *)malloc(sizeof(int64_t) * (size_t)(K + 1)); memcpy(g->ic[j], icols[j], sizeof(int64_t) * (size_t)K);
  }
  for (int j = 0; j < n_fp; ++j) {
    if (borrow) { g->fc[j] = (double *)fcols[j]; continue; }
    g->fc[j] = (double *)malloc(sizeof(double) * (size_t)(K + 1)); memcpy(g->fc[j], fcols[j], sizeof(double) * (size_t)K);
  }
  g->kind = (int *)calloc((size_t)n_nodes, sizeof(int));
  g->c1 = (int *)calloc((size_t)n_nodes, sizeof(int));
  g->c2 = (int *)calloc((size_t)n_nodes, sizeof(int));
  for (int i = 0; i < n_nodes; ++i) {
    int op = nodes[i].op;
    if (op == OP_VAR) g->kind[i] = KD_VAR;
    else if (op <= OP_PAR) g->kind[i] = KD_CONST;
    else if (op >= OP_ADD && op <= OP_POW) {
      int a1 = g->kind[nodes[i].a] != KD_CONST, a2 = g->kind[nodes[i].b] != KD_CONST;
      if (a1 && a2) { g->kind[i] = KD_BIN; g->c1[i] = nodes[i].a; g->c2[i] = nodes[i].b; }
      else if (a1 || a2) { g->kind[i] = KD_UN; g->c1[i] = a1 ? nodes[i].a : nodes[i].b; }
    } else {
      if (g->kind[nodes[i].a] != KD_CONST) { g->kind[i] = KD_UN; g->c1[i] = nodes[i].a; }
    }
  }
  plist l1 = {0}, l2 = {0};
  sym_first(g, n_nodes - 1, &l1);
  sym_hr0(g, n_nodes - 1, &l2);
  int *dummy;
  g->nocc1 = l1.n; g->nocc2 = l2.n;
  g->o1step = compress(&l1, &g->comp1, &g->slot1_idx, &dummy); free(dummy);
  g->o2step = compress(&l2, &g->comp2, &g->slot2_i, &g->slot2_j);
  free(l1.a); free(l1.b); free(l2.a); free(l2.b);
  if (m->row_sorted && !is_obj && g->o1step > 1) {
    /* policy 2, by brute force: order the slots by their column at k = 0 and keep that order only if it is strictly
     * increasing at EVERY support (then each Jacobian row is contiguous and column-sorted: COO array == CSR array) */
    int ns = g->o1step, ok = 1;
    int *pos = (int *)malloc(sizeof(int) * (size_t)ns);          /* pos[p] = old slot at sorted position p */
    for (int s = 0; s < ns; ++s) pos[s] = s;
    for (int a = 1; a < ns; ++a) {                                 /* insertion sort, stable */
      int v = pos[a], b = a;
      while (b > 0 && idx_val(g, g->slot1_idx[pos[b - 1]], 0) > idx_val(g, g->slot1_idx[v], 0)) { pos[b] = pos[b - 1]; --b; }
      pos[b] = v;
    }
    for (int64_t k = 0; k < K && ok; ++k)
      for (int p = 0; p + 1 < ns; ++p)
        if (idx_val(g, g->slot1_idx[pos[p]], k) >= idx_val(g, g->slot1_idx[pos[p + 1]], k)) { ok = 0; break; }
    if (ok) {
      int *inv = (int *)malloc(sizeof(int) * (size_t)ns), *si = (int *)malloc(sizeof(int) * (size_t)(ns + 1));
      for (int p = 0; p < ns; ++p) { inv[pos[p]] = p; si[p] = g->slot1_idx[pos[p]]; }
      for (int i = 0; i < g->nocc1; ++i) g->comp1[i] = inv[g->comp1[i]];
      memcpy(g->slot1_idx, si, sizeof(int) * (size_t)ns);
      free(inv); free(si);
    } else m->all_sorted = -1;
    free(pos);
  }
  m->ngen++;
  return m->ngen - 1;
}

int orc_add_gen(omodel *m, int is_obj, const onode *nodes, int n_nodes, const oindex *idx, int n_idx,
                int64_t K, int n_int, const int64_t *const *icols, int n_fp, const double *const *fcols,
                double lcon, double ucon) {
  return add_gen(m, is_obj, nodes, n_nodes, idx, n_idx, K, n_int, icols, n_fp, fcols, lcon, ucon, 0);
}
/* same, but the iterator columns stay owned by the caller (generators over the same iterator share them) */
int orc_add_gen_borrow(omodel *m, int is_obj, const onode *nodes, int n_nodes, const oindex *idx, int n_idx,
                       int64_t K, int n_int, const int64_t *const *icols, int n_fp, const double *const *fcols,
                       double lcon, double ucon) {
  return add_gen(m, is_obj, nodes, n_nodes, idx, n_idx, K, n_int, icols, n_fp, fcols, lcon, ucon, 1);
}

/* offsets (nlp.jl restatement): rows and Jacobian slots over constraint generators in emission
 * order; Hessian slots over ALL objective generators first, then constraint generators */
void orc_finalize(omodel *m) {
  m->ncon = m->nnzj = m->nnzh = 0;
  for (int i = 0; i < m->ngen; ++i) { ogen *g = &m->g[i]; if (g->is_obj) { g->o2 = m->nnzh; m->nnzh += g->K * g->o2step; } }
  for (int i = 0; i < m->ngen; ++i) {
    ogen *g = &m->g[i];
    if (g->is_obj) continue;
    g->o0 = m->ncon; m->ncon += g->K;
    g->o1 = m->nnzj; m->nnzj += g->K * g->o1step;
    g->o2 = m->nnzh; m->nnzh += g->K * g->o2step;
  }
}
int64_t orc_ncon(const omodel *m) { return m->ncon; }
int64_t orc_nnzj(const omodel *m) { return m->nnzj; }
int64_t orc_nnzh(const omodel *m) { return m->nnzh; }
void orc_gen_info(const omodel *m, int i, int64_t *out) {
  const ogen *g = &m->g[i];
  out[0] = g->o0; out[1] = g->o1; out[2] = g->o2; out[3] = g->o1step; out[4] = g->o2step; out[5] = g->K;
  out[6] = g->nocc1; out[7] = g->nocc2;
}

/* ---- numeric tree: value + local partials per node (AdjointNode / SecondAdjointNode) ------ */
typedef struct { double x, y1, y2, h11, h12, h22; } anode;

static double deg2rad(double d) { return d * (M_PI / 180.0); }

/* f, f', f'' of the unary operators of src/operators.jl:8-43 */
static void unary(int op, double u, double *f, double *d, double *h) {
  double s, c, t, q;
  switch (op) {
    case OP_NEG: *f = -u; *d = -1; *h = 0; break;
    case OP_POS: *f = u; *d = 1; *h = 0; break;
    case OP_INV: *f = 1 / u; *d = -1 / (u * u); *h = 2 / (u * u * u); break;
    case OP_SQRT: s = sqrt(u); *f = s; *d = 1 / (2 * s); *h = -1 / (4 * s * u); break;
    case OP_CBRT: s = cbrt(u); *f = s; *d = 1 / (3 * s * s); *h = -2 / (9 * s * s * u); break;
    case OP_ABS: *f = fabs(u); *d = (u >= 0 ? 1.0 : -1.0); *h = 0; break;
    case OP_ABS2: *f = u * u; *d = 2 * u; *h = 2; break;
    case OP_EXP: s = exp(u); *f = s; *d = s; *h = s; break;
    case OP_EXP2: s = exp2(u); *f = s; *d = s * log(2.0); *h = s * log(2.0) * log(2.0); break;
    case OP_LOG: *f = log(u); *d = 1 / u; *h = -1 / (u * u); break;
    case OP_LOG2: *f = log2(u); *d = 1 / (u * log(2.0)); *h = -1 / (u * u * log(2.0)); break;
    case OP_LOG10: *f = log10(u); *d = 1 / (u * log(10.0)); *h = -1 / (u * u * log(10.0)); break;
    case OP_LOG1P: *f = log1p(u); *d = 1 / (1 + u); *h = -1 / ((1 + u) * (1 + u)); break;
    case OP_SIN: *f = sin(u); *d = cos(u); *h = -sin(u); break;
    case OP_COS: *f = cos(u); *d = -sin(u); *h = -cos(u); break;
    case OP_TAN: t = tan(u); c = cos(u); *f = t; *d = 1 / (c * c); *h = 2 * t / (c * c); break;
    case OP_ASIN: q = 1 - u * u; *f = asin(u); *d = 1 / sqrt(q); *h = u / (q * sqrt(q)); break;
    case OP_ACOS: q = 1 - u * u; *f = acos(u); *d = -1 / sqrt(q); *h = -u / (q * sqrt(q)); break;
    case OP_CSC: s = sin(u); c = cos(u); *f = 1 / s; *d = -c / (s * s); *h = (1 + c * c) / (s * s * s); break;
    case OP_SEC: s = sin(u); c = cos(u); *f = 1 / c; *d = s / (c * c); *h = (1 + s * s) / (c * c * c); break;
    case OP_COT: s = sin(u); c = cos(u); *f = c / s; *d = -1 / (s * s); *h = 2 * c / (s * s * s); break;
    case OP_ATAN: q = 1 + u * u; *f = atan(u); *d = 1 / q; *h = -2 * u / (q * q); break;
    case OP_ACOT: q = 1 + u * u; *f = atan(1 / u); *d = -1 / q; *h = 2 * u / (q * q); break;
    case OP_SINH: *f = sinh(u); *d = cosh(u); *h = sinh(u); break;
    case OP_COSH: *f = cosh(u); *d = sinh(u); *h = cosh(u); break;
    case OP_TANH: t = tanh(u); *f = t; *d = 1 - t * t; *h = -2 * t * (1 - t * t); break;
    case OP_CSCH: s = sinh(u); c = cosh(u); *f = 1 / s; *d = -c / (s * s); *h = (1 + c * c) / (s * s * s); break;
    case OP_SECH: s = sinh(u); c = cosh(u); *f = 1 / c; *d = -s / (c * c); *h = (s * s - 1) / (c * c * c); break;
    case OP_COTH: s = sinh(u); c = cosh(u); *f = c / s; *d = -1 / (s * s); *h = 2 * c / (s * s * s); break;
    case OP_ATANH: q = 1 - u * u; *f = atanh(u); *d = 1 / q; *h = 2 * u / (q * q); break;
    case OP_ACOTH: q = 1 - u * u; *f = atanh(1 / u); *d = 1 / q; *h = 2 * u / (q * q); break;
    case OP_SIND: case OP_COSD: case OP_TAND: case OP_CSCD: case OP_SECD: case OP_COTD: {
      static const int base[6] = {OP_SIN, OP_COS, OP_TAN, OP_CSC, OP_SEC, OP_COT};
      double k = M_PI / 180.0;
      unary(base[op - OP_SIND], deg2rad(u), f, d, h);
      *d *= k; *h *= k * k; break;
    }
    case OP_ATAND: case OP_ACOTD: {
      double k = 180.0 / M_PI;
      unary(op == OP_ATAND ? OP_ATAN : OP_ACOT, u, f, d, h);
      *f *= k; *d *= k; *h *= k; break;
    }
    default: *f = *d = *h = NAN;
  }
}

/* evaluate the whole tape at support k into t[] (forward sweep building the adjoint tree) */
static void forward(const omodel *m, const ogen *g, int64_t k, const double *x, anode *t, int second) {
  for (int i = 0; i < g->n; ++i) {
    const onode *nd = &g->nd[i];
    anode *o = &t[i];
    o->y1 = o->y2 = o->h11 = o->h12 = o->h22 = 0;
    int op = nd->op;
    if (op == OP_CONST) { o->x = nd->c; continue; }
    if (op == OP_FIELD) { o->x = g->fc[nd->a][k]; continue; }
    if (op == OP_VAR) { o->x = x[idx_val(g, nd->a, k) - 1]; continue; }
    if (op == OP_PAR) { o->x = m->theta[idx_val(g, nd->a, k) - 1]; continue; }
    if (op >= OP_NEG) { unary(op, t[nd->a].x, &o->x, &o->y1, &o->h11); continue; }
    double a = t[nd->a].x, b = t[nd->b].x;
    int a1 = g->kind[nd->a] != KD_CONST, a2 = g->kind[nd->b] != KD_CONST;
    switch (op) {
      case OP_ADD: o->x = a + b; o->y1 = 1; o->y2 = 1; break;
      case OP_SUB: o->x = a - b; o->y1 = 1; o->y2 = -1; break;
      case OP_MUL: o->x = a * b; o->y1 = b; o->y2 = a; o->h12 = 1; break;
      case OP_DIV: o->x = a / b; o->y1 = 1 / b; o->y2 = -a / (b * b); o->h12 = -1 / (b * b); o->h22 = 2 * a / (b * b * b); break;
      case OP_POW:
        o->x = (b == 2.0) ? a * a : pow(a, b);
        if (a1) { o->y1 = (b == 2.0) ? 2 * a : b * pow(a, b - 1); o->h11 = (b == 2.0) ? 2.0 : b * (b - 1) * pow(a, b - 2); }
        if (a2) { double la = log(a); o->y2 = o->x * la; o->h22 = o->x * la * la; }
        if (a1 && a2) { double la = log(a); o->h12 = pow(a, b - 1) * (1 + b * la); }
        break;
    }
    /* a constant operand turns the node into a unary one acting on the active child */
    if (g->kind[i] == KD_UN) {
      if (!a1) { o->y1 = o->y2; o->h11 = o->h22; }
      o->y2 = o->h12 = o->h22 = 0;
    }
    (void)second;
  }
}

/* ---- first-order reverse passes ------------------------------------------------------------ */
typedef struct {
  const ogen *g; const anode *t; int64_t k; int cnt;
  double *vals;            /* jac_coord / sparse gradient: vals[o1 + o1step*k + comp(cnt)] += adj */
  int64_t off;
  double *dense;           /* grad!: dense[var] += adj                                          */
  const double *v; double acc; double *jtv; double vrow; /* jprod / jtprod                      */
  int64_t *rows, *cols; int64_t row;                       /* structure                         */
  int mode;                /* 0 coord, 1 dense grad, 2 structure, 3 jprod, 4 jtprod            */
} fctx;

static void jrpass(fctx *c, int t, double adj) {
  const ogen *g = c->g;
  switch (g->kind[t]) {
    case KD_VAR: {
      int64_t var = idx_val(g, g->nd[t].a, c->k);
      int slot = g->comp1[c->cnt++];
      switch (c->mode) {
        case 0: c->vals[c->off + slot] += adj; break;
        case 1: c->dense[var - 1] += adj; break;
        case 2: c->rows[c->off + slot] = c->row; c->cols[c->off + slot] = var; break;
        case 3: c->acc += adj * c->v[var - 1]; break;
        case 4: c->jtv[var - 1] += adj * c->vrow; break;
      }
      break;
    }
    case KD_UN: jrpass(c, g->c1[t], adj * c->t[t].y1); break;
    case KD_BIN:
      if (!g->order) { jrpass(c, g->c1[t], adj * c->t[t].y1); jrpass(c, g->c2[t], adj * c->t[t].y2); }
      else { jrpass(c, g->c2[t], adj * c->t[t].y2); jrpass(c, g->c1[t], adj * c->t[t].y1); }
      break;
    default: break;
  }
}

/* ---- second-order reverse passes ----------------------------------------------------------- */
typedef struct {
  const ogen *g; const anode *t; int64_t k; int cnt;
  double *vals; int64_t off;
  int64_t *rows, *cols;
  const double *v; double *hv;
  int mode; /* 0 coord, 2 structure, 3 hprod */
} sctx;

static void emit2(sctx *c, int t1, int t2, double val, int diag_leaf) {
  const ogen *g = c->g;
  int64_t i = idx_val(g, g->nd[t1].a, c->k), j = idx_val(g, g->nd[t2].a, c->k);
  int slot = g->comp2[c->cnt++];
  if (!diag_leaf && i == j) val = 2 * val; /* cross pair that lands on the diagonal */
  switch (c->mode) {
    case 0: c->vals[c->off + slot] += val; break;
    case 2: c->rows[c->off + slot] = i >= j ? i : j; c->cols[c->off + slot] = i >= j ? j : i; break;
    case 3:
      if (i == j) c->hv[i - 1] += val * c->v[i - 1];
      else { c->hv[i - 1] += val * c->v[j - 1]; c->hv[j - 1] += val * c->v[i - 1]; }
      break;
  }
}
/* local partial w.r.t. the i-th child IN VISITING ORDER (kids()); a tape may name the same node as both children */
static double kid_partial(const ogen *g, const anode *T, int t, int i) {
  if (g->kind[t] == KD_VAR) return 1.0;
  if (g->kind[t] == KD_UN) return T[t].y1;
  return ((i == 0) != (g->order != 0)) ? T[t].y1 : T[t].y2;
}
static void hdrpass(sctx *c, int t1, int t2, double adj) {
  const ogen *g = c->g; const anode *T = c->t;
  if (g->kind[t1] == KD_VAR && g->kind[t2] == KD_VAR) { emit2(c, t1, t2, adj, 0); return; }
  int a[2], b[2], na = kids(g, t1, a), nb = kids(g, t2, b);
  for (int i = 0; i < na; ++i)
    for (int j = 0; j < nb; ++j) {
      double v = adj;
      if (g->kind[t1] != KD_VAR) v *= kid_partial(g, T, t1, i);
      if (g->kind[t2] != KD_VAR) v *= kid_partial(g, T, t2, j);
      hdrpass(c, a[i], b[j], v);
    }
}
static void hrpass(sctx *c, int t, double adj, double adj2) {
  const ogen *g = c->g; const anode *n = &c->t[t];
  switch (g->kind[t]) {
    case KD_VAR: emit2(c, t, t, adj2, 1); break;
    case KD_UN: hrpass(c, g->c1[t], adj * n->y1, adj2 * (n->y1 * n->y1) + adj * n->h11); break;
    case KD_BIN: {
      double cross = adj2 * n->y1 * n->y2 + adj * n->h12;
      if (!g->order) {
        hrpass(c, g->c1[t], adj * n->y1, adj2 * (n->y1 * n->y1) + adj * n->h11);
        hrpass(c, g->c2[t], adj * n->y2, adj2 * (n->y2 * n->y2) + adj * n->h22);
        hdrpass(c, g->c1[t], g->c2[t], cross);
      } else {
        hrpass(c, g->c2[t], adj * n->y2, adj2 * (n->y2 * n->y2) + adj * n->h22);
        hrpass(c, g->c1[t], adj * n->y1, adj2 * (n->y1 * n->y1) + adj * n->h11);
        hdrpass(c, g->c2[t], g->c1[t], cross);
      }
      break;
    }
    default: break;
  }
}
/* top-level pass: NO second-order adjoint (Appendix A.4: "adj2 == 0, no cross term") — through +, -, const*subtree only
 * the first-order adjoint travels; the first nonlinear node hands its children adj*h (not 0*y*y + adj*h) */
static void hrpass0(sctx *c, int t, double adj) {
  const ogen *g = c->g; const anode *n = &c->t[t];
  if (g->kind[t] == KD_VAR || g->kind[t] == KD_CONST) return;
  if (passthrough(g, t)) {
    int op = g->nd[t].op;
    if (g->kind[t] == KD_BIN) {
      if (!g->order) { hrpass0(c, g->c1[t], adj); hrpass0(c, g->c2[t], op == OP_SUB ? -adj : adj); }
      else { hrpass0(c, g->c2[t], op == OP_SUB ? -adj : adj); hrpass0(c, g->c1[t], adj); }
    } else if (op == OP_MUL) {
      hrpass0(c, g->c1[t], adj * n->y1);
    } else {
      hrpass0(c, g->c1[t], n->y1 < 0 ? -adj : adj);
    }
    return;
  }
  if (g->kind[t] == KD_UN) {
    hrpass(c, g->c1[t], adj * n->y1, adj * n->h11);
  } else if (!g->order) {
    hrpass(c, g->c1[t], adj * n->y1, adj * n->h11);
    hrpass(c, g->c2[t], adj * n->y2, adj * n->h22);
    hdrpass(c, g->c1[t], g->c2[t], adj * n->h12);
  } else {
    hrpass(c, g->c2[t], adj * n->y2, adj * n->h22);
    hrpass(c, g->c1[t], adj * n->y1, adj * n->h11);
    hdrpass(c, g->c2[t], g->c1[t], adj * n->h12);
  }
}

/* ---- NLPModels-level entry points ----------------------------------------------------------- */
#define FOR_GEN(m, g, want_obj) for (int _i = 0; _i < (m)->ngen; ++_i) if (((g) = &(m)->g[_i])->is_obj == (want_obj))

double orc_obj(const omodel *m, const double *x) {
  double f = 0; const ogen *g;
  FOR_GEN(m, g, 1) {
    double fg = 0;
#pragma omp parallel
    {
      anode *t = (anode *)malloc(sizeof(anode) * (size_t)g->n);
#pragma omp for reduction(+ : fg)
      for (int64_t k = 0; k < g->K; ++k) { forward(m, g, k, x, t, 0); fg += t[g->n - 1].x; }
      free(t);
    }
    f += fg;
  }
  return f;
}

void orc_cons(const omodel *m, const double *x, double *cv) {
  const ogen *g;
  for (int64_t i = 0; i < m->ncon; ++i) cv[i] = 0;
  FOR_GEN(m, g, 0) {
#pragma omp parallel
    {
      anode *t = (anode *)malloc(sizeof(anode) * (size_t)g->n);
#pragma omp for
      for (int64_t k = 0; k < g->K; ++k) { forward(m, g, k, x, t, 0); cv[g->o0 + k] += t[g->n - 1].x; }
      free(t);
    }
  }
}

void orc_grad(const omodel *m, const double *x, double *gr) {
  const ogen *g;
  for (int64_t i = 0; i < m->nvar; ++i) gr[i] = 0;
  FOR_GEN(m, g, 1) { /* sequential: dense scatter-add */
    anode *t = (anode *)malloc(sizeof(anode) * (size_t)g->n);
    for (int64_t k = 0; k < g->K; ++k) {
      forward(m, g, k, x, t, 0);
      fctx c = {g, t, k, 0, 0, 0, gr, 0, 0, 0, 0, 0, 0, 0, 1};
      jrpass(&c, g->n - 1, 1.0);
    }
    free(t);
  }
}

void orc_jac_structure(const omodel *m, int64_t *rows, int64_t *cols) {
  const ogen *g;
  FOR_GEN(m, g, 0) {
    anode *t = (anode *)calloc((size_t)g->n, sizeof(anode)); /* structure needs no values */
    for (int64_t k = 0; k < g->K; ++k) {
      fctx c = {g, t, k, 0, 0, g->o1 + g->o1step * k, 0, 0, 0, 0, 0, rows, cols, g->o0 + k + 1, 2};
      jrpass(&c, g->n - 1, 1.0);
    }
    free(t);
  }
}

void orc_jac_coord(const omodel *m, const double *x, double *vals) {
  const ogen *g;
  for (int64_t i = 0; i < m->nnzj; ++i) vals[i] = 0;
  FOR_GEN(m, g, 0) {
#pragma omp parallel
    {
      anode *t = (anode *)malloc(sizeof(anode) * (size_t)g->n);
#pragma omp for
      for (int64_t k = 0; k < g->K; ++k) {
        forward(m, g, k, x, t, 0);
        fctx c = {g, t, k, 0, vals, g->o1 + g->o1step * k, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        jrpass(&c, g->n - 1, 1.0);
      }
      free(t);
    }
  }
}

void orc_jprod(const omodel *m, const double *x, const double *v, double *jv) {
  const ogen *g;
  for (int64_t i = 0; i < m->ncon; ++i) jv[i] = 0;
  FOR_GEN(m, g, 0) {
#pragma omp parallel
    {
      anode *t = (anode *)malloc(sizeof(anode) * (size_t)g->n);
#pragma omp for
      for (int64_t k = 0; k < g->K; ++k) { /* rows are independent: the sum of one row stays sequential */
        forward(m, g, k, x, t, 0);
        fctx c = {g, t, k, 0, 0, 0, 0, v, 0, 0, 0, 0, 0, 0, 3};
        jrpass(&c, g->n - 1, 1.0);
        jv[g->o0 + k] += c.acc;
      }
      free(t);
    }
  }
}

void orc_jtprod(const omodel *m, const double *x, const double *v, double *jtv) {
  const ogen *g;
  for (int64_t i = 0; i < m->nvar; ++i) jtv[i] = 0;
  FOR_GEN(m, g, 0) {
    anode *t = (anode *)malloc(sizeof(anode) * (size_t)g->n);
    for (int64_t k = 0; k < g->K; ++k) {
      forward(m, g, k, x, t, 0);
      fctx c = {g, t, k, 0, 0, 0, 0, 0, 0, jtv, v[g->o0 + k], 0, 0, 0, 4};
      jrpass(&c, g->n - 1, 1.0);
    }
    free(t);
  }
}

void orc_hess_structure(const omodel *m, int64_t *rows, int64_t *cols) {
  for (int pass = 1; pass >= 0; --pass) {
    const ogen *g;
    FOR_GEN(m, g, pass) {
      anode *t = (anode *)calloc((size_t)g->n, sizeof(anode)); /* structure needs no values */
      for (int64_t k = 0; k < g->K; ++k) {
        sctx c = {g, t, k, 0, 0, g->o2 + g->o2step * k, rows, cols, 0, 0, 2};
        hrpass0(&c, g->n - 1, 0.0);
      }
      free(t);
    }
  }
}

void orc_hess_coord(const omodel *m, const double *x, const double *y, double sigma, double *vals) {
  for (int64_t i = 0; i < m->nnzh; ++i) vals[i] = 0;
  for (int pass = 1; pass >= 0; --pass) {
    const ogen *g;
    FOR_GEN(m, g, pass) {
      if (!g->is_obj && !y) continue;
#pragma omp parallel
      {
        anode *t = (anode *)malloc(sizeof(anode) * (size_t)g->n);
#pragma omp for
        for (int64_t k = 0; k < g->K; ++k) {
          forward(m, g, k, x, t, 1);
          sctx c = {g, t, k, 0, vals, g->o2 + g->o2step * k, 0, 0, 0, 0, 0};
          hrpass0(&c, g->n - 1, g->is_obj ? sigma : y[g->o0 + k]);
        }
        free(t);
      }
    }
  }
}

void orc_hprod(const omodel *m, const double *x, const double *y, const double *v, double sigma, double *hv) {
  for (int64_t i = 0; i < m->nvar; ++i) hv[i] = 0;
  for (int pass = 1; pass >= 0; --pass) {
    const ogen *g;
    FOR_GEN(m, g, pass) {
      if (!g->is_obj && !y) continue;
      anode *t = (anode *)malloc(sizeof(anode) * (size_t)g->n);
      for (int64_t k = 0; k < g->K; ++k) {
        forward(m, g, k, x, t, 1);
        sctx c = {g, t, k, 0, 0, 0, 0, 0, v, hv, 3};
        hrpass0(&c, g->n - 1, g->is_obj ? sigma : y[g->o0 + k]);
      }
      free(t);
    }
  }
}

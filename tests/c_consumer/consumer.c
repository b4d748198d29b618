/* consumer.c — a plain-C consumer of the C ABI (include/iexa.h), test infrastructure.
 *
 * What a host in ANY language does at the drop-in boundary (the Julia glue does the same through ccall):
 * dlopen libiexa_b200.so, describe a model as SoA iterator columns + postfix tapes, finalize, call the NLPModels callbacks
 * with plain host pointers.  No Python, no torch, no C++ on this side.  The model is small enough for closed forms:
 *
 *   min   sum_k w_k (x_k - t_k)^2                                   k = 1..K   (one objective generator; t, w: fp fields)
 *   s.t.  x_k y_k - theta_1 sin(x_k)            = 0                 k = 1..K   (generator over the same iterator; a PAR leaf)
 *         y_{j+1} - y_j - 0.25 x_{j+1}          = 0                 j = 1..K-1 (shifted indices: idx +- const, transform.jl:485-505)
 *
 * Every result is compared with the closed form after summing the COO triplets into dense matrices, so nothing here depends
 * on the slot order inside a generator.  usage: consumer <path/to/libiexa_b200.so> [nodevice]
 * `nodevice`: IEXA_F_NO_DEVICE plan — meta works, every evaluation call must FAIL with a message (no CPU fallback).          */
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/iexa.h"

#define K 7
#define NV (2 * K)
#define NC (2 * K - 1)

static void *lib;
#define FN(name) __typeof__(&name) p_##name = (__typeof__(&name))dlsym(lib, #name); if (!p_##name) { fprintf(stderr, "missing symbol %s\n", #name); return 2; }
#define CHECK(call) do { int rc_ = (call); if (rc_ != 0) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, p_iexa_last_error()); return 1; } } while (0)

static int close_to(double a, double b) { return fabs(a - b) <= 1e-14 + 1e-12 * fmax(fabs(a), fabs(b)); }

int main(int argc, char **argv) {
  if (argc < 2) { fprintf(stderr, "usage: consumer <libiexa_b200.so> [nodevice]\n"); return 2; }
  const int nodevice = argc > 2 && strcmp(argv[2], "nodevice") == 0;
  lib = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
  FN(iexa_last_error) FN(iexa_plan_create) FN(iexa_plan_destroy) FN(iexa_add_var) FN(iexa_add_par) FN(iexa_itr_base)
  FN(iexa_add_con) FN(iexa_add_obj) FN(iexa_finalize) FN(iexa_get_meta) FN(iexa_jac_structure) FN(iexa_hess_structure)
  FN(iexa_obj) FN(iexa_grad) FN(iexa_cons) FN(iexa_jac_coord) FN(iexa_hess_coord) FN(iexa_jprod) FN(iexa_jtprod) FN(iexa_hprod)
  FN(iexa_set_par) FN(iexa_get_vector)

  iexa_plan *p = NULL;
  CHECK(p_iexa_plan_create(&p, 1));
  double x0[K], lo[K], up[K];
  for (int k = 0; k < K; ++k) { x0[k] = 0.3 + 0.1 * k; lo[k] = -5.0; up[k] = 5.0; }
  int64_t xoff = -1, yoff = -1, poff = -1;
  CHECK(p_iexa_add_var(p, K, x0, lo, up, &xoff));
  CHECK(p_iexa_add_var(p, K, NULL, NULL, NULL, &yoff));
  const double theta1 = 2.0;
  CHECK(p_iexa_add_par(p, 1, &theta1, &poff));
  if (xoff != 0 || yoff != K || poff != 0) { fprintf(stderr, "offsets %lld %lld %lld\n", (long long)xoff, (long long)yoff, (long long)poff); return 1; }

  /* iterators: SoA columns */
  int64_t ik[K], ij[K - 1];
  double t[K], w[K];
  for (int k = 0; k < K; ++k) { ik[k] = k + 1; t[k] = 0.25 * k; w[k] = 0.5 + 0.125 * k; }
  for (int j = 0; j < K - 1; ++j) ij[j] = j + 1;
  const int64_t *ic1[1] = {ik}, *ic2[1] = {ij};
  const double *fc1[2] = {t, w};
  int32_t it1 = -1, it2 = -1;
  CHECK(p_iexa_itr_base(p, K, 1, ic1, 2, fc1, &it1));
  CHECK(p_iexa_itr_base(p, K - 1, 1, ic2, 0, NULL, &it2));

  /* objective  w * abs2(x[i] - t) */
  iexa_index oi[1]; memset(oi, 0, sizeof oi);
  oi[0].base = xoff; oi[0].nterms = 1; oi[0].col[0] = 0; oi[0].coef[0] = 1;
  iexa_node on[6]; memset(on, 0, sizeof on);
  on[0].op = IEXA_OP_FIELD; on[0].a = 1;
  on[1].op = IEXA_OP_VAR; on[1].a = 0;
  on[2].op = IEXA_OP_FIELD; on[2].a = 0;
  on[3].op = IEXA_OP_SUB; on[3].a = 1; on[3].b = 2;
  on[4].op = IEXA_OP_ABS2; on[4].a = 3;
  on[5].op = IEXA_OP_MUL; on[5].a = 0; on[5].b = 4;
  CHECK(p_iexa_add_obj(p, on, 6, oi, 1, it1));

  /* rows 1..K:  x[i]*y[i] - theta[1]*sin(x[i]) */
  iexa_index ci[3]; memset(ci, 0, sizeof ci);
  ci[0].base = xoff; ci[0].nterms = 1; ci[0].coef[0] = 1;
  ci[1].base = yoff; ci[1].nterms = 1; ci[1].coef[0] = 1;
  ci[2].base = poff + 1; ci[2].nterms = 0;
  iexa_node cn[8]; memset(cn, 0, sizeof cn);
  cn[0].op = IEXA_OP_VAR; cn[0].a = 0;
  cn[1].op = IEXA_OP_VAR; cn[1].a = 1;
  cn[2].op = IEXA_OP_MUL; cn[2].a = 0; cn[2].b = 1;
  cn[3].op = IEXA_OP_PAR; cn[3].a = 2;
  cn[4].op = IEXA_OP_VAR; cn[4].a = 0;
  cn[5].op = IEXA_OP_SIN; cn[5].a = 4;
  cn[6].op = IEXA_OP_MUL; cn[6].a = 3; cn[6].b = 5;
  cn[7].op = IEXA_OP_SUB; cn[7].a = 2; cn[7].b = 6;
  int64_t row1 = -1, row2 = -1;
  CHECK(p_iexa_add_con(p, cn, 8, ci, 3, it1, 0.0, 0.0, &row1));

  /* rows K+1..2K-1:  y[j+1] - y[j] - 0.25*x[j+1] */
  iexa_index di[3]; memset(di, 0, sizeof di);
  di[0].base = yoff + 1; di[0].nterms = 1; di[0].coef[0] = 1;
  di[1].base = yoff; di[1].nterms = 1; di[1].coef[0] = 1;
  di[2].base = xoff + 1; di[2].nterms = 1; di[2].coef[0] = 1;
  iexa_node dn[7]; memset(dn, 0, sizeof dn);
  dn[0].op = IEXA_OP_VAR; dn[0].a = 0;
  dn[1].op = IEXA_OP_VAR; dn[1].a = 1;
  dn[2].op = IEXA_OP_SUB; dn[2].a = 0; dn[2].b = 1;
  dn[3].op = IEXA_OP_CONST; dn[3].c = 0.25;
  dn[4].op = IEXA_OP_VAR; dn[4].a = 2;
  dn[5].op = IEXA_OP_MUL; dn[5].a = 3; dn[5].b = 4;
  dn[6].op = IEXA_OP_SUB; dn[6].a = 2; dn[6].b = 5;
  CHECK(p_iexa_add_con(p, dn, 7, di, 3, it2, 0.0, 0.0, &row2));
  if (row1 != 0 || row2 != K) { fprintf(stderr, "row offsets %lld %lld\n", (long long)row1, (long long)row2); return 1; }

  CHECK(p_iexa_finalize(p, 0, 0, 1, nodevice ? IEXA_F_NO_DEVICE : IEXA_F_DEFAULT));
  iexa_meta m;
  CHECK(p_iexa_get_meta(p, &m));
  if (m.nvar != NV || m.ncon != NC || m.npar != 1 || m.nnzj != 2 * K + 3 * (K - 1) || m.nnzh <= 0) {
    fprintf(stderr, "meta: nvar %lld ncon %lld npar %lld nnzj %lld nnzh %lld\n", (long long)m.nvar, (long long)m.ncon, (long long)m.npar, (long long)m.nnzj, (long long)m.nnzh);
    return 1;
  }
  double xs[NV];
  CHECK(p_iexa_get_vector(p, 0, xs));
  for (int k = 0; k < K; ++k) if (xs[k] != x0[k] || xs[K + k] != 0.0) { fprintf(stderr, "x0 mismatch\n"); return 1; }

  double x[NV], y[NC], v[NV], u[NC];
  for (int i = 0; i < NV; ++i) { x[i] = 0.2 + 0.07 * i - 0.003 * i * i; v[i] = cos(1.0 + i); }
  for (int r = 0; r < NC; ++r) { y[r] = sin(0.5 + r); u[r] = cos(2.0 * r); }
  const double sigma = 0.7;
  double c[NC];
  if (nodevice) {
    /* no device engine: the evaluation entry points must fail loudly — there is no CPU fallback behind this ABI */
    int rc = p_iexa_cons(p, x, c, IEXA_MEM_HOST, NULL);
    const char *msg = p_iexa_last_error();
    if (rc == 0 || !msg || !msg[0]) { fprintf(stderr, "iexa_cons succeeded without a device engine\n"); return 1; }
    double f = 0.0;
    if (p_iexa_obj(p, x, &f, IEXA_MEM_HOST, NULL) == 0) { fprintf(stderr, "iexa_obj succeeded without a device engine\n"); return 1; }
    printf("OK nodevice: nvar=%lld ncon=%lld nnzj=%lld nnzh=%lld; evaluation refused: %s\n", (long long)m.nvar, (long long)m.ncon, (long long)m.nnzj, (long long)m.nnzh, msg);
    p_iexa_plan_destroy(p);
    return 0;
  }

  /* ---- closed forms ---- */
  double th = theta1, f_ref = 0.0, g_ref[NV], c_ref[NC], J[NC][NV], H[NV][NV];
  int bad = 0;
  for (int pass = 0; pass < 2; ++pass) {   /* second pass: after set_parameter! */
    memset(g_ref, 0, sizeof g_ref); memset(J, 0, sizeof J); memset(H, 0, sizeof H); f_ref = 0.0;
    for (int k = 0; k < K; ++k) {
      const double xk = x[k], yk = x[K + k];
      f_ref += w[k] * (xk - t[k]) * (xk - t[k]);
      g_ref[k] = 2.0 * w[k] * (xk - t[k]);
      c_ref[k] = xk * yk - th * sin(xk);
      J[k][k] = yk - th * cos(xk);
      J[k][K + k] = xk;
      H[k][k] += sigma * 2.0 * w[k] + y[k] * th * sin(xk);
      H[K + k][k] += y[k];
    }
    for (int j = 0; j < K - 1; ++j) {
      c_ref[K + j] = x[K + j + 1] - x[K + j] - 0.25 * x[j + 1];
      J[K + j][K + j + 1] = 1.0; J[K + j][K + j] = -1.0; J[K + j][j + 1] = -0.25;
    }
    double f = 0.0, g[NV];
    CHECK(p_iexa_obj(p, x, &f, IEXA_MEM_HOST, NULL));
    CHECK(p_iexa_grad(p, x, g, IEXA_MEM_HOST, NULL));
    CHECK(p_iexa_cons(p, x, c, IEXA_MEM_HOST, NULL));
    if (!close_to(f, f_ref)) { fprintf(stderr, "obj %.17g vs %.17g\n", f, f_ref); ++bad; }
    for (int i = 0; i < NV; ++i) if (!close_to(g[i], g_ref[i])) { fprintf(stderr, "grad[%d] %.17g vs %.17g\n", i, g[i], g_ref[i]); ++bad; }
    for (int r = 0; r < NC; ++r) if (!close_to(c[r], c_ref[r])) { fprintf(stderr, "cons[%d] %.17g vs %.17g\n", r, c[r], c_ref[r]); ++bad; }

    int64_t *jr = malloc(sizeof(int64_t) * m.nnzj), *jc = malloc(sizeof(int64_t) * m.nnzj);
    int32_t *hr = malloc(sizeof(int32_t) * m.nnzh), *hc = malloc(sizeof(int32_t) * m.nnzh);   /* Int32 index buffers work too */
    double *jv = malloc(sizeof(double) * m.nnzj), *hv = malloc(sizeof(double) * m.nnzh);
    CHECK(p_iexa_jac_structure(p, jr, jc, 8, IEXA_MEM_HOST, NULL));
    CHECK(p_iexa_hess_structure(p, hr, hc, 4, IEXA_MEM_HOST, NULL));
    CHECK(p_iexa_jac_coord(p, x, jv, IEXA_MEM_HOST, NULL));
    CHECK(p_iexa_hess_coord(p, x, y, sigma, hv, IEXA_MEM_HOST, NULL));
    double Jd[NC][NV], Hd[NV][NV];
    memset(Jd, 0, sizeof Jd); memset(Hd, 0, sizeof Hd);
    for (int64_t s = 0; s < m.nnzj; ++s) {
      if (jr[s] < 1 || jr[s] > NC || jc[s] < 1 || jc[s] > NV) { fprintf(stderr, "jac structure out of range\n"); return 1; }
      Jd[jr[s] - 1][jc[s] - 1] += jv[s];
    }
    for (int64_t s = 0; s < m.nnzh; ++s) {
      if (hr[s] < hc[s] || hc[s] < 1 || hr[s] > NV) { fprintf(stderr, "hess structure not lower-triangular\n"); return 1; }
      Hd[hr[s] - 1][hc[s] - 1] += hv[s];
    }
    for (int r = 0; r < NC; ++r) for (int i = 0; i < NV; ++i) if (!close_to(Jd[r][i], J[r][i])) { fprintf(stderr, "J[%d][%d] %.17g vs %.17g\n", r, i, Jd[r][i], J[r][i]); ++bad; }
    for (int a = 0; a < NV; ++a) for (int b = 0; b <= a; ++b) if (!close_to(Hd[a][b], H[a][b])) { fprintf(stderr, "H[%d][%d] %.17g vs %.17g\n", a, b, Hd[a][b], H[a][b]); ++bad; }

    /* matrix-free products against the dense matrices */
    double Jv[NC], Jtu[NV], Hv[NV];
    CHECK(p_iexa_jprod(p, x, v, Jv, IEXA_MEM_HOST, NULL));
    CHECK(p_iexa_jtprod(p, x, u, Jtu, IEXA_MEM_HOST, NULL));
    CHECK(p_iexa_hprod(p, x, y, v, sigma, Hv, IEXA_MEM_HOST, NULL));
    for (int r = 0; r < NC; ++r) { double s = 0; for (int i = 0; i < NV; ++i) s += J[r][i] * v[i]; if (fabs(s - Jv[r]) > 1e-13 * (1 + fabs(s))) { fprintf(stderr, "jprod[%d]\n", r); ++bad; } }
    for (int i = 0; i < NV; ++i) { double s = 0; for (int r = 0; r < NC; ++r) s += J[r][i] * u[r]; if (fabs(s - Jtu[i]) > 1e-13 * (1 + fabs(s))) { fprintf(stderr, "jtprod[%d]\n", i); ++bad; } }
    for (int a = 0; a < NV; ++a) {
      double s = 0;
      for (int b = 0; b < NV; ++b) s += (a >= b ? H[a][b] : H[b][a]) * v[b];
      if (fabs(s - Hv[a]) > 1e-13 * (1 + fabs(s))) { fprintf(stderr, "hprod[%d] %.17g vs %.17g\n", a, Hv[a], s); ++bad; }
    }
    free(jr); free(jc); free(hr); free(hc); free(jv); free(hv);
    if (pass == 0) { th = -1.25; CHECK(p_iexa_set_par(p, poff, 1, &th)); }   /* set_parameter!: theta changes, the plan does not */
  }
  CHECK(p_iexa_plan_destroy(p));
  if (bad) { fprintf(stderr, "%d mismatches\n", bad); return 1; }
  printf("OK: obj, grad, cons, jac, hess, jprod, jtprod, hprod through the C ABI from plain C match the closed forms (before and after set_par)\n");
  return 0;
}

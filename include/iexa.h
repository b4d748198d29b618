/*
 * iexa.h — C ABI of the B200-native NLP evaluation engine ("iexa").
 *
 * This is the drop-in boundary for the hot path of infiniteopt/InfiniteExaModels.jl:
 * the NLPModels callbacks (obj, grad!, cons!, jac_structure!/jac_coord!,
 * hess_structure!/hess_coord!, jprod!/jtprod!/hprod!) that MadNLP / Ipopt call on the
 * model that `ExaTranscriptionBackend` builds.  The reference has NO C ABI of its own
 * (it is pure Julia and reaches the evaluator through ExaModels.jl); every entry point
 * below names the reference call site (file:line under /root/reference) it replaces.
 *
 * Conventions
 *   - every function returns an int32 status (IEXA_OK == 0); iexa_last_error() gives a
 *     thread-local message.  No C++ exception and no exit() crosses this boundary.
 *   - indices handed to the caller (rows/cols, offsets) are 1-based like the Julia
 *     solvers expect; index expressions in tapes are 1-based as in transform.jl.
 *   - `memspace` says where the caller's buffers live (IEXA_MEM_HOST / IEXA_MEM_DEVICE).
 *     Host buffers are copied inside the call (see iexa_host_register for page-locking).
 *   - `stream` is a cudaStream_t cast to void* (NULL = legacy default stream).  Device
 *     calls are asynchronous on that stream except where a scalar is returned (iexa_obj).
 *   - the engine never keeps caller pointers after a call returns.
 *   - NaN/Inf are propagated untouched (the reference never throws during evaluation;
 *     MadNLP maps them to INVALID_NUMBER_*, ext/InfiniteExaModelsMadNLP.jl:78-87).
 *   - there is NO CPU fallback: evaluation entry points fail with IEXA_ERR_CUDA when no
 *     CUDA device / kernel image is available.
 */
#ifndef IEXA_H
#define IEXA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IEXA_VERSION 100 /* 0.1.0 */

/* ---- status codes ---------------------------------------------------------------- */
enum {
  IEXA_OK = 0,
  IEXA_ERR_INVALID = 1,     /* bad argument / malformed tape                          */
  IEXA_ERR_STATE = 2,       /* call not allowed in this plan state (e.g. before finalize) */
  IEXA_ERR_CUDA = 3,        /* CUDA runtime/driver error, or no device                 */
  IEXA_ERR_NVRTC = 4,       /* run-time specialisation failed                          */
  IEXA_ERR_UNSUPPORTED = 5, /* operator outside src/operators.jl:2-46                  */
  IEXA_ERR_NOMEM = 6
};

enum {
  IEXA_MEM_HOST = 0,
  IEXA_MEM_DEVICE = 1,
  /* host buffers, and x is bit-identical to the x of the previous host-memory call on this plan: the engine
   * reuses the device copy of x it already holds instead of uploading it again.  This is Ipopt's `new_x == false`
   * (every Ipopt callback — eval_g, eval_jac_g, eval_h, … — receives that flag; NLPModelsIpopt drops it, the binding in
   * INTEGRATION.md forwards it).  Falls back to a plain upload when the engine holds no copy yet. */
  IEXA_MEM_HOST_SAME_X = 2
};

/* ---- tape operators ---------------------------------------------------------------
 * One entry per operator of the reference's table src/operators.jl:2-46 plus the four
 * leaf kinds that src/transform.jl:290-330 (_map_variable) can produce.               */
enum {
  /* leaves */
  IEXA_OP_CONST = 0, /* literal c                      (transform.jl:340, Null :393)  */
  IEXA_OP_FIELD = 1, /* fp iterator field  data_src[alias]       (transform.jl:320-322) */
  IEXA_OP_VAR = 2,   /* x[index expr]      (transform.jl:290-319)                     */
  IEXA_OP_PAR = 3,   /* theta[index expr]  (transform.jl:323-330)                     */
  /* binary  (operators.jl:3-7) */
  IEXA_OP_ADD = 10, IEXA_OP_SUB = 11, IEXA_OP_MUL = 12, IEXA_OP_DIV = 13, IEXA_OP_POW = 14,
  /* unary   (operators.jl:8-43; NEG/POS are the unary forms of :- and :+) */
  IEXA_OP_NEG = 20, IEXA_OP_POS = 21, IEXA_OP_INV = 22, IEXA_OP_SQRT = 23, IEXA_OP_CBRT = 24,
  IEXA_OP_ABS = 25, IEXA_OP_ABS2 = 26, IEXA_OP_EXP = 27, IEXA_OP_EXP2 = 28, IEXA_OP_LOG = 29,
  IEXA_OP_LOG2 = 30, IEXA_OP_LOG10 = 31, IEXA_OP_LOG1P = 32, IEXA_OP_SIN = 33, IEXA_OP_COS = 34,
  IEXA_OP_TAN = 35, IEXA_OP_ASIN = 36, IEXA_OP_ACOS = 37, IEXA_OP_CSC = 38, IEXA_OP_SEC = 39,
  IEXA_OP_COT = 40, IEXA_OP_ATAN = 41, IEXA_OP_ACOT = 42, IEXA_OP_SIND = 43, IEXA_OP_COSD = 44,
  IEXA_OP_TAND = 45, IEXA_OP_CSCD = 46, IEXA_OP_SECD = 47, IEXA_OP_COTD = 48, IEXA_OP_ATAND = 49,
  IEXA_OP_ACOTD = 50, IEXA_OP_SINH = 51, IEXA_OP_COSH = 52, IEXA_OP_TANH = 53,
  IEXA_OP_CSCH = 54, /* the TRUE csch; the reference maps :csch to csc (operators.jl:41) —
                        a lowering that wants bug-compatibility emits IEXA_OP_CSC instead */
  IEXA_OP_SECH = 55, IEXA_OP_COTH = 56, IEXA_OP_ATANH = 57, IEXA_OP_ACOTH = 58,
  IEXA_OP__END = 59
};

/* One node of a postfix tape: children precede parents, the root is the last node.
 * A tape is a TREE (a node id that is referenced twice is evaluated as two subtrees,
 * like the expression objects ExaModels builds).                                      */
typedef struct iexa_node {
  int32_t op; /* IEXA_OP_*                                                             */
  int32_t a;  /* unary/binary: lhs child node id | VAR/PAR: index-expr id | FIELD: fp column */
  int32_t b;  /* binary: rhs child node id, otherwise 0                                */
  int32_t pad;
  double c;   /* CONST: the literal                                                    */
} iexa_node;

#define IEXA_MAX_INDEX_TERMS 4
/* Affine integer index expression over the integer columns of the generator's iterator:
 *   idx(k) = base + sum_j coef[j] * int_col[col[j]](k)            (1-based result)
 * This is what transform.jl's Variable[data_src[group_alias]...] (:309-310,:316-318),
 * `idx ± const` in make_reduced_expr (:485-505) and the (i1,i2) collocation pairs
 * (:593-596) reduce to under column-major addressing.                                 */
typedef struct iexa_index {
  int64_t base;
  int32_t nterms;
  int32_t col[IEXA_MAX_INDEX_TERMS];
  int32_t pad;
  int64_t coef[IEXA_MAX_INDEX_TERMS];
} iexa_index;

typedef struct iexa_plan iexa_plan;

typedef struct iexa_meta {
  int64_t nvar, ncon, npar, nobj_gen, ncon_gen;
  int64_t nnzj, nnzh, nnzg; /* global COO sizes (nnzg: sparse objective-gradient slots) */
  int64_t loc_ncon, loc_nnzj, loc_nnzh; /* what THIS rank owns (== global when world==1) */
  int32_t minimize, rank, world, device;
  int32_t n_kernels_specialised; /* NVRTC-specialised kernel images in use             */
  int32_t pad;
} iexa_meta;

/* finalize flags */
enum {
  IEXA_F_DEFAULT = 0,
  IEXA_F_NO_SPECIALISE = 1, /* use only the AOT tape-interpreter kernels (no NVRTC)    */
  IEXA_F_NO_DEVICE = 2      /* compile the plan on the host only (structure queries,
                               byte accounting); every evaluation call then FAILS       */
};

const char *iexa_last_error(void);
int32_t iexa_version(void);

/* ---- plan construction (replaces ExaModels.ExaCore(...) at transform.jl:815 and the
 *      builder calls listed below) ------------------------------------------------- */
int32_t iexa_plan_create(iexa_plan **out, int32_t minimize);
int32_t iexa_plan_destroy(iexa_plan *p);

/* Plan options — set BEFORE the first iexa_add_con / iexa_add_obj.
 *   IEXA_OPT_SLOT_ORDER: the order in which the symbolic reverse passes meet the Var leaves decides which COO slot a
 *     variable (pair) gets inside a generator.  ExaModels 0.11.2 (Project.toml:23; reached from transform.jl:458,559,597)
 *     is absent from the reference tree, so its order is a hypothesis — it is therefore DATA here, not
 *     code: LEFT_TO_RIGHT (default; children inner1 then inner2) or RIGHT_TO_LEFT (inner2 then inner1).  A dump of the
 *     real package (julia/dump_golden.jl) only has to name the policy; kernels do not change.
 *   IEXA_OPT_STRICT_IEEE: 1 (DEFAULT) = structural zeros are multiplied at run time (0*x, 0/x are not folded unless x is a
 *     literal), so NaN / Inf propagate through the derivative formulas exactly as in a run-time AD that multiplies the
 *     numbers (tests/test_nan_semantics.py: NaN pattern identical to the oracle on poisoned inputs for every BASELINE
 *     config); 0 folds them: identical results for finite inputs, a subset of the NaNs otherwise, 1.5-2.5 % faster.  */
enum { IEXA_OPT_SLOT_ORDER = 1, IEXA_OPT_STRICT_IEEE = 2 };
/*     JAC_ROW_SORTED (an engine-native layout, NOT a hypothesis about ExaModels): LEFT_TO_RIGHT for everything except the
 *     first-order slots of constraint generators, which are put in increasing column order whenever that order is the
 *     same at every support.  A generator's rows are contiguous in the COO array (ExaModels' own layout:
 *     slot = o1 + o1step*(k-1) + c), so the array jac_coord! writes then IS a CSR value array — iexa_jac_csr_rowptr gives the
 *     row pointers, iexa_jac_structure's cols are the column indices, and the COO->CSR pass (iexa_csr_apply) that a
 *     KKT assembly otherwise pays per iteration disappears.  Same nnz, same values, a permutation of the default order. */
enum { IEXA_SLOT_ORDER_LEFT_TO_RIGHT = 0, IEXA_SLOT_ORDER_RIGHT_TO_LEFT = 1, IEXA_SLOT_ORDER_JAC_ROW_SORTED = 2 };
int32_t iexa_set_option(iexa_plan *p, int32_t key, int64_t value);

/* ExaModels.add_var  — transform.jl:113 (finite), :154 (infinite + derivative vars).
 * x0/lvar/uvar may be NULL (0, -inf, +inf).  offset_out: 0-based offset of the block. */
int32_t iexa_add_var(iexa_plan *p, int64_t n, const double *x0, const double *lvar,
                     const double *uvar, int64_t *offset_out);
/* ExaModels.add_par  — transform.jl:127 (finite parameters), :179 (parameter functions) */
int32_t iexa_add_par(iexa_plan *p, int64_t n, const double *vals, int64_t *offset_out);
/* element-wise patch of x0/lvar/uvar before finalize — transform.jl:216-231
 * which: 0 = x0, 1 = lvar, 2 = uvar; i is a 1-based variable index                    */
int32_t iexa_patch_var(iexa_plan *p, int32_t which, int64_t i, double value);

/* Iterators.  A base iterator is the SoA form of transform.jl:31's Vector{NamedTuple}:
 * n_int integer columns (group_idx / i1,i2 / support indices) and n_fp fp64 columns
 * (support values, d_arg coefficients, quadrature weight c), each of length K.
 * Integer columns equal to 1..K are detected and never stored or loaded.              */
int32_t iexa_itr_base(iexa_plan *p, int64_t K, int32_t n_int, const int64_t *const *int_cols,
                      int32_t n_fp, const double *const *fp_cols, int32_t *itr_out);
/* ---- device-side transcription (src/transform.jl:2-38 iterators, :161-183 parameter functions, :618-633 measure
 *      coefficients): columns described by a closed form are GENERATED on the device by a kernel — no K-long host array, no
 *      upload — with numpy's / InfiniteOpt's arithmetic restated operation by operation, so the values are bit-identical to
 *      the host path's.  int_cols[j] == NULL is the column 1..K.  gens[j].kind: 0 = data (fp_cols[j]);
 *      IEXA_GEN_LINSPACE       v(j) = a + j*(b-a)/(n-1), v(n-1) = b, K == n        (supports of an independent parameter)
 *      IEXA_GEN_LINSPACE_MID   that grid interleaved with its interval midpoints, K == 2n-1 (OrthogonalCollocation(3) nodes)
 *      IEXA_GEN_TRAPEZOID      trapezoid weights of the earlier fp column `src` of this iterator (default integral)
 *      IEXA_GEN_CONST          v(j) = a                                                                       */
enum { IEXA_GEN_DATA = 0, IEXA_GEN_LINSPACE = 1, IEXA_GEN_LINSPACE_MID = 2, IEXA_GEN_TRAPEZOID = 3, IEXA_GEN_CONST = 4 };
typedef struct iexa_colgen {
  int32_t kind, src;
  int64_t n;
  double a, b;
} iexa_colgen;
int32_t iexa_itr_generated(iexa_plan *p, int64_t K, int32_t n_int, const int64_t *const *int_cols, int32_t n_fp,
                           const iexa_colgen *gens, const double *const *fp_cols, int32_t *itr_out);
/* ExaModels.add_par for a parameter FUNCTION (transform.jl:161-183) given as a tape over an iterator: the block of theta
 * (one entry per support, iterator order) is evaluated on the device at iexa_finalize.  Leaves: CONST, FIELD, PAR.   */
int32_t iexa_add_par_function(iexa_plan *p, const iexa_node *nodes, int32_t n_nodes, const iexa_index *idx,
                              int32_t n_idx, int32_t itr, int64_t *offset_out);
/* host copy of fp column `col` of base iterator `itr` as the DEVICE holds it (generated columns: downloaded)      */
int32_t iexa_debug_get_column(iexa_plan *p, int32_t itr, int32_t col, double *out_host);

/* Product iterator, first factor fastest (transform.jl:445, :541, :591, :670): columns are
 * the concatenation of the factors' columns; nothing is materialised.                 */
int32_t iexa_itr_product(iexa_plan *p, int32_t n, const int32_t *itrs, int32_t *itr_out);

/* ExaModels.add_con — transform.jl:458 (constraints), :559 (derivative approximations),
 * :597 (collocation restrictions).  row_offset_out: 0-based first row.                */
int32_t iexa_add_con(iexa_plan *p, const iexa_node *nodes, int32_t n_nodes,
                     const iexa_index *idx, int32_t n_idx, int32_t itr, double lcon,
                     double ucon, int64_t *row_offset_out);
/* ExaModels.add_obj — transform.jl:614, :700, :741                                    */
int32_t iexa_add_obj(iexa_plan *p, const iexa_node *nodes, int32_t n_nodes,
                     const iexa_index *idx, int32_t n_idx, int32_t itr);

/* ExaModels.ExaModel(core) — infiniteopt_backend.jl:156.  Compiles every generator
 * (symbolic sparsity + slot layout + AD programs), uploads columns/theta, selects or
 * specialises kernels.  rank/world shard the support ranges (world==1: whole model).  */
int32_t iexa_finalize(iexa_plan *p, int32_t device, int32_t rank, int32_t world, uint32_t flags);

int32_t iexa_get_meta(const iexa_plan *p, iexa_meta *out);
/* host copies of the NLPModelMeta vectors (get_x0/get_y0: infiniteopt_backend.jl:600-601)
 * which: 0 x0, 1 lvar, 2 uvar (length nvar); 3 lcon, 4 ucon, 5 y0 (length ncon)        */
int32_t iexa_get_vector(const iexa_plan *p, int32_t which, double *out);
int32_t iexa_set_vector(iexa_plan *p, int32_t which, const double *in); /* x0 / y0 warm start */

/* ExaModels.set_parameter! — infiniteopt_backend.jl:522,:546 ; model.θ — :479          */
/* iexa_set_par has no stream argument: it synchronises the DEVICE before it overwrites theta, so it is ordered against
 * callbacks in flight on any stream (CUDA.jl task streams, torch side streams are non-blocking with respect to the
 * legacy default stream).  iexa_set_par_stream is the asynchronous form: the update is enqueued on `stream` like a
 * callback (values are staged in a pinned buffer of the engine; vals_host may be reused as soon as the call returns). */
int32_t iexa_set_par(iexa_plan *p, int64_t offset0, int64_t n, const double *vals_host);
int32_t iexa_set_par_stream(iexa_plan *p, int64_t offset0, int64_t n, const double *vals_host, void *stream);
/* world > 1: a rank's device holds the slices of theta its own supports read (iexa_device_bytes); iexa_set_par writes the resident
 * part of the range and skips the rest.  iexa_get_par serves the host mirror: complete for blocks given as values; for a block
 * evaluated on the device (iexa_add_par_function) a rank of world > 1 has computed — and returns — only the entries its supports
 * read (zeros elsewhere; IEXA_NO_PFUNC_SHARDING=1 evaluates whole blocks on every rank).                                      */
int32_t iexa_get_par(const iexa_plan *p, int64_t offset0, int64_t n, double *vals_host);

/* ---- NLPModels callbacks (ExaModels 0.11.2 methods reached from
 *      ext/InfiniteExaModelsMadNLP.jl:49-50,64 and ext/InfiniteExaModelsIpopt.jl:48-49,59-60)
 * idx_bytes: 4 or 8 (Int32 / Int64 index buffers).  Buffers sized by the LOCAL counts
 * of iexa_meta when world > 1.                                                         */
int32_t iexa_jac_structure(iexa_plan *p, void *rows, void *cols, int32_t idx_bytes,
                           int32_t memspace, void *stream);
int32_t iexa_hess_structure(iexa_plan *p, void *rows, void *cols, int32_t idx_bytes,
                            int32_t memspace, void *stream);
/* is_csr_out = 1 when the jac_coord! array is already in CSR order (rows contiguous, columns strictly increasing inside a
 * row): policy IEXA_SLOT_ORDER_JAC_ROW_SORTED and every constraint generator had a static column order.
 * iexa_jac_csr_rowptr then writes the ncon+1 (local counts when world > 1) 0-based row pointers; IEXA_ERR_STATE otherwise. */
int32_t iexa_jac_is_csr(const iexa_plan *p, int32_t *is_csr_out);
int32_t iexa_jac_csr_rowptr(iexa_plan *p, void *rowptr, int32_t idx_bytes, int32_t memspace, void *stream);
int32_t iexa_obj(iexa_plan *p, const double *x, double *f_host, int32_t memspace, void *stream);
/* world > 1: g is written only inside this rank's iexa_x_ranges (its own supports, shared variables, halos); the rank contributes
 * nothing elsewhere and leaves those entries UNTOUCHED (zero g first if the whole vector is to be all-reduced)              */
int32_t iexa_grad(iexa_plan *p, const double *x, double *g, int32_t memspace, void *stream);
int32_t iexa_cons(iexa_plan *p, const double *x, double *c, int32_t memspace, void *stream);
int32_t iexa_jac_coord(iexa_plan *p, const double *x, double *vals, int32_t memspace, void *stream);
/* y may be NULL (objective-only Hessian)                                               */
int32_t iexa_hess_coord(iexa_plan *p, const double *x, const double *y, double obj_weight,
                        double *vals, int32_t memspace, void *stream);
/* cons! + jac_coord! + hess_coord! at the SAME (x, y) in one call — what an interior-point iteration asks for (MadNLP's
 * eval_f / eval_cons / eval_jac / eval_lag_hess wrappers all run at the current iterate).  Device buffers: ONE fused kernel —
 * every constraint group evaluates value, first and second order from one program (x / theta / columns loaded once, sin /
 * cos of a state once) — instead of three launches with three fill / drain phases; results are identical to the three
 * callbacks.  Host buffers, or when the fused kernel is unavailable: the three callbacks in sequence (x uploaded once). */
int32_t iexa_eval3(iexa_plan *p, const double *x, const double *y, double obj_weight, double *c, double *jac_vals,
                   double *hess_vals, int32_t memspace, void *stream);
/* Matrix-free products (NLPModels jprod! / jtprod! / hprod!: the model handed to the solver at
 * ext/InfiniteExaModelsMadNLP.jl:49-50 / ext/InfiniteExaModelsIpopt.jl:48-49 carries the full API).  Fused kernels:
 * the first / second order programs with a product epilogue — no COO values are written or read.
 *   jprod:  Jv[row] = sum_c d1_c * v[col_c] in registers, one coalesced store per row, no atomics: bit-reproducible.
 *   jtprod / hprod: one sum per touched variable and support in registers; single-writer blocks are plain stores,
 *   the rest atomics (warp-shuffle reduced for shared variables) in a second, ordered launch.
 * v of jprod / hprod and the results of jtprod / hprod have nvar entries; v of jtprod and Jv have the LOCAL row count.
 * world > 1: jtprod / hprod return this rank's partial sums (all-reduce the vector).  The kernels of these three entry
 * points are compiled on the first call of any of them.                                                          */
int32_t iexa_jprod(iexa_plan *p, const double *x, const double *v, double *Jv, int32_t memspace,
                   void *stream);
int32_t iexa_jtprod(iexa_plan *p, const double *x, const double *v, double *Jtv,
                    int32_t memspace, void *stream);
int32_t iexa_hprod(iexa_plan *p, const double *x, const double *y, const double *v,
                   double obj_weight, double *Hv, int32_t memspace, void *stream);
/* device-side objective partial (no host sync): writes 1 double to f_dev               */
int32_t iexa_obj_device(iexa_plan *p, const double *x_dev, double *f_dev, void *stream);

/* Optional page-locking of caller-owned HOST vectors that are reused on every iteration (the
 * Ipopt-style path): host<->device copies then run at PCIe speed.  Explicit because only the
 * caller knows the buffers' lifetime; unregister before freeing them.                       */
int32_t iexa_host_register(iexa_plan *p, void *buf, int64_t bytes);
int32_t iexa_host_unregister(iexa_plan *p, void *buf);

/* ---- sharding queries (world > 1).  A segment maps a contiguous local range of rows /
 *      Jacobian slots / Hessian slots to its position in the global (unsharded) arrays. */
typedef struct iexa_segment {
  int64_t global_start, local_start, length; /* 0-based */
} iexa_segment;
/* which: 0 rows, 1 jac slots, 2 hess slots.  Returns the count; fills up to cap.       */
int64_t iexa_segments(const iexa_plan *p, int32_t which, iexa_segment *out, int64_t cap);
/* Device memory the ENGINE holds for this plan on this rank, in bytes:
 *   out6[0] iterator columns resident here        out6[1] the same columns of the unsharded model
 *   out6[2] theta resident here                   out6[3] 8 * npar
 *   out6[4] programs, descriptors, work tables    out6[5] staging buffers of the host-memory path (grown on first use)
 * world > 1: a rank keeps the slice of every column that its own supports visit, and theta — addressed by global parameter
 * indices inside the kernels — is a full-length virtual address range in which only the granules (2 MB) covering the rank's
 * slices are backed by memory (CUDA virtual-memory API; IEXA_NO_VMM=1 / IEXA_NO_COLUMN_SLICES=1 keep everything whole).
 * x, y and the outputs belong to the caller: x is full-length by contract (global indices), of which a rank touches iexa_x_ranges. */
int32_t iexa_device_bytes(const iexa_plan *p, int64_t *out6);
/* variable indices (1-based) whose gradient entries receive contributions from more than
 * one rank (finite / shared variables and shard-boundary halos): the slice that must be
 * all-reduced after iexa_grad.  Returns the count; fills up to cap.                    */
int64_t iexa_shared_vars(const iexa_plan *p, int64_t *out, int64_t cap);
/* the same set as merged, sorted ranges (global_start == local_start, 0-based).  Covers shifted references at the shard
 * boundaries ((y[i+1]-y[i])^2: the boundary entry is written by both neighbours), indices of product / restricted
 * iterators (w[s]*z[t]^2 over (t, s): every z[t] is written by every rank) and constant indices.  Conservative: an
 * entry may be listed although one rank contributes to it, never the other way round.                          */
int64_t iexa_shared_ranges(const iexa_plan *p, iexa_segment *out, int64_t cap);
/* the parts of x this rank's callbacks READ: its own supports of every variable block, the replicated finite /
 * shared variables and the halo of shifted references (y[i-1] of finite differences, the lower-bound / internal
 * nodes of collocation elements) at the shard boundaries — merged, sorted ranges (global_start == local_start,
 * 0-based; a conservative cover).  A distributed solver keeps only these ranges of x current on this rank
 * (halo exchange instead of a broadcast of the whole iterate).  Returns the count; fills up to cap.            */
int64_t iexa_x_ranges(const iexa_plan *p, iexa_segment *out, int64_t cap);

/* bytes of x a host-memory callback uploads on this rank: 8*nvar when world == 1, else the iexa_x_ranges only      */
int64_t iexa_host_x_bytes(const iexa_plan *p);

/* ---- byte accounting used by bench.py's roofline: ALGORITHMIC bytes of
 *      one call of each callback, computed from the finalised plan.
 * which: 0 obj, 1 grad, 2 cons, 3 jac_coord, 4 hess_coord, 5 jprod, 6 jtprod, 7 hprod, 8 eval3 (inputs once + c + both value arrays)
 * (products: the inputs their first / second order programs load, the touched part of v — or y and v — and every
 * entry of the dense result once; no COO values are materialised)                          */
int64_t iexa_algorithmic_bytes(const iexa_plan *p, int32_t which);
/* number of kernel launches one call of callback `which` performs                      */
int32_t iexa_launches_per_call(const iexa_plan *p, int32_t which);

/* why run-time specialisation is not in use for this plan ("" when it is): NVRTC/driver missing,
 * compile error, or the generated source exceeded the compile budget                      */
const char *iexa_engine_note(const iexa_plan *p);

/* ---- diagnostics: the CUDA translation unit that iexa_finalize specialises with NVRTC.
 *      _source returns its length (copies up to cap-1 bytes); _compile runs NVRTC for sm_100a
 *      without loading the image (works on a machine without a GPU).                       */
/* cap < 0 (capacity -cap) and *cubin_bytes == -1 on entry select the second translation unit: the jprod! / jtprod! /
 * hprod! kernels, which the engine compiles on the first product call                                   */
int64_t iexa_debug_codegen_source(const iexa_plan *p, char *buf, int64_t cap);
/* regroup a host-only plan with (1) / without (0) shape canonicalisation before inspecting its source */
int32_t iexa_debug_set_class_mode(iexa_plan *p, int32_t on);
int32_t iexa_debug_codegen_compile(const iexa_plan *p, int64_t *cubin_bytes);
/* the same for a named kernel set: 0 = the five callbacks, 1 = jprod! / jtprod! / hprod!, 2 = the fused iexa_eval3 kernel */
int64_t iexa_debug_codegen_source_of(const iexa_plan *p, int32_t set, char *buf, int64_t cap);
int32_t iexa_debug_codegen_compile_of(const iexa_plan *p, int32_t set, int64_t *cubin_bytes);
/* compiled images are cached per process AND on disk ($IEXA_CACHE_DIR | $XDG_CACHE_HOME/iexa_b200 | ~/.cache/iexa_b200;
 * IEXA_CACHE_DIR=off disables), keyed by the generated source: counts of NVRTC compilations / disk hits of this process */
int32_t iexa_debug_cache_stats(int32_t *nvrtc_compiles, int32_t *disk_hits);

/* ---- COO -> CSR value permutation feeding cuDSS (today MadNLPGPU's transfer! kernel,
 *      caller side of ext/InfiniteExaModelsMadNLP.jl:49-50).  Setup sorts the 1-based
 *      COO pattern once; apply sums duplicates into CSR order with no atomics.         */
typedef struct iexa_csr iexa_csr;
int32_t iexa_csr_create(iexa_csr **out, int64_t nrows, int64_t ncols, int64_t nnz,
                        const void *rows, const void *cols, int32_t idx_bytes,
                        int32_t memspace, int32_t device);
/* Same, with a locality key per COO entry (entries with nearby keys are evaluated from nearby supports): the
 * apply pass then walks the CSR entries grouped by key, so that every consumer of a COO tile runs while the
 * tile is in cache.  iexa_coo_locality fills the keys of the Jacobian (which = 0) / Hessian (which = 1) slots of a
 * plan (relative position of the slot's 128-support bucket in its generator, scaled to [0, 2^20)); a caller that
 * concatenates COO blocks into a KKT pattern (MadNLP) concatenates the keys the same way and gives any
 * constant to the entries it adds itself.  keys == NULL is iexa_csr_create.                             */
int32_t iexa_csr_create_keyed(iexa_csr **out, int64_t nrows, int64_t ncols, int64_t nnz,
                              const void *rows, const void *cols, int32_t idx_bytes,
                              const int32_t *keys, int32_t memspace, int32_t device);
int32_t iexa_coo_locality(iexa_plan *p, int32_t which, int32_t *keys, int32_t memspace, void *stream);
int32_t iexa_csr_destroy(iexa_csr *h);
int64_t iexa_csr_nnz(const iexa_csr *h);
/* rowptr: nrows+1 entries, colind: csr_nnz entries, 0-based int32 (cuDSS convention)    */
int32_t iexa_csr_pattern(const iexa_csr *h, int32_t *rowptr, int32_t *colind, int32_t memspace);
int32_t iexa_csr_apply(iexa_csr *h, const double *coo_vals, double *csr_vals, int32_t memspace,
                       void *stream);

/* ---- x halo exchange + small all-reduce over NVLink peer memory (one process per GPU, CUDA IPC) — csrc/halo.cu.
 *      A sharded solver keeps on every rank only the part of the iterate it owns; before the callbacks of a new iterate
 *      each rank PUSHES the shared variables and the shard-boundary halos it owns straight into its readers' x buffers
 *      (peer stores) — one small kernel per rank, flags in peer memory, no host synchronisation, no staging copies.
 *      Which entries go where follows from iexa_x_ranges of all ranks (dist.py: x_partition).  Replaces the NCCL
 *      point-to-point exchange (0.12 ms for 180 doubles on 8 GPUs); sharded analogue of the x every callback of
 *      ext/InfiniteExaModelsMadNLP.jl:49-50 receives.
 *      Setup: create; export the handles of MY x buffer and flag block; ship them to every peer (torch.distributed);
 *      connect each peer's handles; set_sends / set_recvs.  exchange and allreduce_small are COLLECTIVE (every rank,
 *      same order) and asynchronous on `stream`.  Waits inside the kernels are bounded (~4 s): iexa_halo_status != 0
 *      reports an expired wait instead of a hung GPU.                                                              */
typedef struct iexa_halo iexa_halo;
int32_t iexa_halo_create(iexa_halo **out, int32_t device, int32_t rank, int32_t world);
int32_t iexa_halo_export(iexa_halo *h, const void *x_dev, unsigned char *handle_x /*64 B*/, int64_t *x_offset,
                         unsigned char *handle_flags /*64 B*/);
int32_t iexa_halo_connect(iexa_halo *h, int32_t peer, const unsigned char *handle_x, int64_t x_offset,
                          const unsigned char *handle_flags);
/* lo_hi: n_ranges pairs [lo, hi), 0-based global positions of x that this rank owns and `peer` reads              */
int32_t iexa_halo_set_sends(iexa_halo *h, int32_t peer, int64_t n_ranges, const int64_t *lo_hi);
int32_t iexa_halo_set_recvs(iexa_halo *h, int32_t n_peers, const int32_t *peers);
int32_t iexa_halo_exchange(iexa_halo *h, const double *x_dev, void *stream);
/* buf[0..n) <- sum over ranks, n <= 1024, in place; summed in rank order: deterministic, bit-identical on all ranks */
int32_t iexa_halo_allreduce_small(iexa_halo *h, double *buf_dev, int32_t n, void *stream);
int64_t iexa_halo_status(iexa_halo *h);
int32_t iexa_halo_destroy(iexa_halo *h);

#ifdef __cplusplus
}
#endif
#endif /* IEXA_H */

// dag.hpp — hash-consed scalar expression DAG used by the plan compiler.
//
// The reference stack (ExaModels 0.11.2, reached from src/transform.jl:458,559,597,614,700)
// differentiates every generator at RUN time by building AdjointNode / SecondAdjointNode
// trees per support point and walking them recursively.  This engine differentiates each
// generator ONCE, at plan-compile time, into a DAG of primitive fp64 operations whose
// outputs are the constraint value, the o1step Jacobian slot values and the o2step
// Hessian slot values.  The DAG is then scheduled into a register program that the
// sm_100a kernels (tape interpreter, or an NVRTC-specialised image of the same program)
// execute once per support point.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <unordered_map>
#include <vector>

namespace iexa {

enum DOp : int32_t {
  D_CONST = 0, // c
  D_FIELD,     // fp iterator column a
  D_LOADX,     // x[index slot a]
  D_LOADP,     // theta[index slot a]
  D_W,         // root weight of group member a: y[row] for constraints, obj_weight for objectives
  D_SEL2,      // (index slot a == index slot b) ? 2.0 : 1.0   (lower-triangle diagonal rule)
  D_CPAR,      // per-instance constant a of a shape class (plan.hpp: Group::is_class); opaque to folding
  D_LOADV,     // v[index slot a]: the vector of a matrix-free product (jprod! / hprod!), addressed like x
  D_SELNE,     // (index slot a != index slot b) ? 1.0 : 0.0   (mirror contribution of a lower-triangle entry)
  D_ADD, D_SUB, D_MUL, D_DIV, D_NEG, D_POW,
  D_SQRT, D_CBRT, D_ABS, D_SIGNP, D_EXP, D_EXP2, D_LOG, D_LOG2, D_LOG10, D_LOG1P,
  D_SIN, D_COS, D_TAN, D_ASIN, D_ACOS, D_ATAN, D_SINH, D_COSH, D_TANH, D_ATANH,
  D_OUT, // program pseudo-instruction: output slot dst <- reg a (never a DAG node)
  D__N
};

inline bool dop_is_unary(int32_t op) { return op == D_NEG || (op >= D_SQRT && op <= D_ATANH); }
inline bool dop_is_binary(int32_t op) { return (op >= D_ADD && op <= D_DIV) || op == D_POW; }

struct DNode {
  int32_t op;
  int32_t a, b;
  double c;
};

inline double host_unary(int32_t op, double x) {
  switch (op) {
    case D_NEG: return -x;
    case D_SQRT: return std::sqrt(x);
    case D_CBRT: return std::cbrt(x);
    case D_ABS: return std::fabs(x);
    case D_SIGNP: return x >= 0 ? 1.0 : -1.0;
    case D_EXP: return std::exp(x);
    case D_EXP2: return std::exp2(x);
    case D_LOG: return std::log(x);
    case D_LOG2: return std::log2(x);
    case D_LOG10: return std::log10(x);
    case D_LOG1P: return std::log1p(x);
    case D_SIN: return std::sin(x);
    case D_COS: return std::cos(x);
    case D_TAN: return std::tan(x);
    case D_ASIN: return std::asin(x);
    case D_ACOS: return std::acos(x);
    case D_ATAN: return std::atan(x);
    case D_SINH: return std::sinh(x);
    case D_COSH: return std::cosh(x);
    case D_TANH: return std::tanh(x);
    case D_ATANH: return std::atanh(x);
    default: return NAN;
  }
}

class Dag {
 public:
  std::vector<DNode> nodes;
  // IEEE-strict mode: 0*x and 0/x are NOT folded to 0 unless x is a literal, so a structural zero times Inf / NaN
  // evaluates to NaN exactly like a run-time AD that multiplies the numbers (ExaModels' reverse passes compute
  // adj2*y^2 + adj*h with literal zeros in y / h).  Everything else that is folded is exact for every input:
  // x+0, x*1, x*(-1), x/1, pow(x,1), pow(x,2) = x*x, and pow(x,0) = 1 (IEEE: also for NaN and Inf).
  bool strict = false;

  int cnst(double c) { return intern(D_CONST, 0, 0, c); }
  int field(int col) { return intern(D_FIELD, col, 0, 0.0); }
  int loadx(int islot) { return intern(D_LOADX, islot, 0, 0.0); }
  int loadp(int islot) { return intern(D_LOADP, islot, 0, 0.0); }
  int w(int member = 0) { return intern(D_W, member, 0, 0.0); }
  int cpar(int j) { return intern(D_CPAR, j, 0, 0.0); }
  int loadv(int islot) { return intern(D_LOADV, islot, 0, 0.0); }
  int selne(int ia, int ib) {
    if (ia == ib) return cnst(0.0);
    if (ia > ib) std::swap(ia, ib);
    return intern(D_SELNE, ia, ib, 0.0);
  }
  int sel2(int ia, int ib) {
    if (ia == ib) return cnst(2.0);
    if (ia > ib) std::swap(ia, ib);
    return intern(D_SEL2, ia, ib, 0.0);
  }

  bool is_const(int n) const { return nodes[n].op == D_CONST; }
  double cval(int n) const { return nodes[n].c; }
  bool is_c(int n, double v) const { return is_const(n) && nodes[n].c == v; }

  int add(int a, int b) {
    if (is_const(a) && is_const(b)) return cnst(cval(a) + cval(b));
    if (is_c(a, 0.0)) return b;
    if (is_c(b, 0.0)) return a;
    if (a > b) std::swap(a, b);
    return intern(D_ADD, a, b, 0.0);
  }
  int sub(int a, int b) {
    if (is_const(a) && is_const(b)) return cnst(cval(a) - cval(b));
    if (is_c(b, 0.0)) return a;
    if (is_c(a, 0.0)) return neg(b);
    return intern(D_SUB, a, b, 0.0);
  }
  int mul(int a, int b) {
    if (is_const(a) && is_const(b)) return cnst(cval(a) * cval(b));
    // exact for finite operands only: kept as a multiplication in strict mode
    if (!strict && (is_c(a, 0.0) || is_c(b, 0.0))) return cnst(0.0);
    if (is_c(a, 1.0)) return b;
    if (is_c(b, 1.0)) return a;
    if (is_c(a, -1.0)) return neg(b);
    if (is_c(b, -1.0)) return neg(a);
    if (a > b) std::swap(a, b);
    return intern(D_MUL, a, b, 0.0);
  }
  int div(int a, int b) {
    if (is_const(a) && is_const(b)) return cnst(cval(a) / cval(b));
    if (is_c(b, 1.0)) return a;
    if (!strict && is_c(a, 0.0)) return cnst(0.0);
    return intern(D_DIV, a, b, 0.0);
  }
  int neg(int a) {
    if (is_const(a)) return cnst(-cval(a));
    if (nodes[a].op == D_NEG) return nodes[a].a;
    return intern(D_NEG, a, 0, 0.0);
  }
  int pow(int a, int b) {
    if (is_const(a) && is_const(b)) return cnst(std::pow(cval(a), cval(b)));
    if (is_c(b, 1.0)) return a;
    if (is_c(b, 2.0)) return mul(a, a);
    if (is_c(b, 0.0)) return cnst(1.0);
    return intern(D_POW, a, b, 0.0);
  }
  int un(int32_t op, int a) {
    if (op == D_NEG) return neg(a);
    if (is_const(a)) return cnst(host_unary(op, cval(a)));
    return intern(op, a, 0, 0.0);
  }
  int sq(int a) { return mul(a, a); }

 private:
  struct Key {
    int32_t op, a, b;
    uint64_t cbits;
    bool operator==(const Key &o) const {
      return op == o.op && a == o.a && b == o.b && cbits == o.cbits;
    }
  };
  struct KeyHash {
    size_t operator()(const Key &k) const {
      uint64_t h = 1469598103934665603ull;
      auto mix = [&](uint64_t v) { h ^= v; h *= 1099511628211ull; h ^= h >> 29; };
      mix((uint64_t)(uint32_t)k.op);
      mix((uint64_t)(uint32_t)k.a);
      mix((uint64_t)(uint32_t)k.b);
      mix(k.cbits);
      return (size_t)h;
    }
  };
  std::unordered_map<Key, int, KeyHash> cse_;

  int intern(int32_t op, int32_t a, int32_t b, double c) {
    Key k{op, a, b, 0};
    std::memcpy(&k.cbits, &c, 8);
    auto it = cse_.find(k);
    if (it != cse_.end()) return it->second;
    int id = (int)nodes.size();
    nodes.push_back(DNode{op, a, b, c});
    cse_.emplace(k, id);
    return id;
  }
};

// ---- register program (what the kernels execute) ------------------------------------
struct Instr {
  int32_t op;  // DOp
  int32_t dst; // destination register | D_OUT: output slot
  int32_t a;   // register | D_CONST: const-pool index | D_FIELD: fp column slot | D_LOAD*: index slot
  int32_t b;   // register | D_SEL2: second index slot
};

struct Program {
  std::vector<Instr> code;
  std::vector<double> cpool;
  int32_t nreg = 0;
  int32_t nout = 0;
  bool uses_w = false;
  int32_t n_flop_nodes = 0; // arithmetic/transcendental instructions (for reporting)
};

} // namespace iexa

// codegen.cpp — prints the fused register programs of a finalised plan as straight-line CUDA and
// compiles them for sm_100a with NVRTC (see codegen.hpp).  libcuda / libnvrtc are dlopen'ed so
// that the shared library still loads (for symbol / host-logic tests) on a machine without a
// driver; on such a machine build() fails and the caller reports it.
//
// Shape of a generated kernel (one per NLPModels callback):
//   * blockIdx.x -> (group, block of IEXA_BLOCK supports) through the work table;
//   * every scalar the body needs that is not data (support range, output offsets, index bases,
//     column table pointers) sits in a __constant__ array and is addressed with literal offsets,
//     so it costs a constant-bank operand, not a global load;
//   * x / theta / column loads are coalesced: thread t of the block owns support kb + t;
//   * Jacobian / Hessian values of a member with ostep > 1 are staged in shared memory as the
//     [support][slot] tile they occupy in the COO array and written out with fully coalesced
//     256-byte warp stores (the per-support AoS layout o + ostep*k + c is dictated by the
//     reference; a direct store would touch 32 sectors per warp instruction).
#include "codegen.hpp"

#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <mutex>
#include <sstream>

#include "engine.hpp"

namespace iexa {

// ---- minimal driver / NVRTC API surface (dlopen) ------------------------------------------
typedef int CUresult;
typedef void *CUmodule;
typedef void *CUfunction;
typedef void *CUstream;
typedef unsigned long long CUdeviceptr;
typedef struct _nvrtcProgram *nvrtcProgram;

struct DynApi {
  void *cuda = nullptr, *nvrtc = nullptr;
  CUresult (*cuInit)(unsigned) = nullptr;
  CUresult (*cuModuleLoadData)(CUmodule *, const void *) = nullptr;
  CUresult (*cuModuleUnload)(CUmodule) = nullptr;
  CUresult (*cuModuleGetFunction)(CUfunction *, CUmodule, const char *) = nullptr;
  CUresult (*cuModuleGetGlobal_v2)(CUdeviceptr *, size_t *, CUmodule, const char *) = nullptr;
  CUresult (*cuMemcpyHtoD_v2)(CUdeviceptr, const void *, size_t) = nullptr;
  CUresult (*cuMemAlloc_v2)(CUdeviceptr *, size_t) = nullptr;
  CUresult (*cuMemFree_v2)(CUdeviceptr) = nullptr;
  CUresult (*cuFuncSetAttribute)(CUfunction, int, int) = nullptr;
  CUresult (*cuLaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                             unsigned, CUstream, void **, void **) = nullptr;
  CUresult (*cuGetErrorString)(CUresult, const char **) = nullptr;
  CUresult (*cuLaunchKernelEx)(const void *, CUfunction, void **, void **) = nullptr; // optional (PDL launches)
  int (*nvrtcCreateProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *,
                            const char *const *) = nullptr;
  int (*nvrtcCompileProgram)(nvrtcProgram, int, const char *const *) = nullptr;
  int (*nvrtcGetCUBINSize)(nvrtcProgram, size_t *) = nullptr;
  int (*nvrtcGetCUBIN)(nvrtcProgram, char *) = nullptr;
  int (*nvrtcGetProgramLogSize)(nvrtcProgram, size_t *) = nullptr;
  int (*nvrtcGetProgramLog)(nvrtcProgram, char *) = nullptr;
  int (*nvrtcDestroyProgram)(nvrtcProgram *) = nullptr;
  int (*nvrtcVersion)(int *, int *) = nullptr; // optional: part of the on-disk cache key

  std::mutex mu; // lazily initialised from any thread that finalizes a plan
  bool load_nvrtc(std::string &err) {
    std::lock_guard<std::mutex> lk(mu);
    if (nvrtc_ok) return true;
    const char *names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"};
    for (const char *n : names) if ((nvrtc = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!nvrtc) { err = "cannot dlopen libnvrtc"; return false; }
#define L(sym) *(void **)(&sym) = dlsym(nvrtc, #sym); if (!sym) { err = "libnvrtc lacks " #sym; return false; }
    L(nvrtcCreateProgram) L(nvrtcCompileProgram) L(nvrtcGetCUBINSize) L(nvrtcGetCUBIN)
    L(nvrtcGetProgramLogSize) L(nvrtcGetProgramLog) L(nvrtcDestroyProgram)
#undef L
    *(void **)(&nvrtcVersion) = dlsym(nvrtc, "nvrtcVersion");
    nvrtc_ok = true;
    return true;
  }
  bool nvrtc_ok = false, cuda_ok = false;
  bool load_cuda(std::string &err) {
    std::lock_guard<std::mutex> lk(mu);
    if (cuda_ok) return true;
    cuda = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!cuda) { err = "cannot dlopen libcuda.so.1 (no NVIDIA driver)"; return false; }
#define L(sym) *(void **)(&sym) = dlsym(cuda, #sym); if (!sym) { err = "libcuda lacks " #sym; return false; }
    L(cuInit) L(cuModuleLoadData) L(cuModuleUnload) L(cuModuleGetFunction) L(cuModuleGetGlobal_v2)
    L(cuMemcpyHtoD_v2) L(cuMemAlloc_v2) L(cuMemFree_v2) L(cuFuncSetAttribute) L(cuLaunchKernel) L(cuGetErrorString)
#undef L
    *(void **)(&cuLaunchKernelEx) = dlsym(cuda, "cuLaunchKernelEx");
    cuda_ok = true;
    return true;
  }
};
static DynApi &api() { static DynApi a; return a; }

// ---- source generation --------------------------------------------------------------------------
static const char *kPrelude = R"CUDA(
// generated by iexa (infiniteexamodels.jl_b200/csrc/codegen.cpp) — do not edit
struct WorkItem { int gen; int blk; };
__device__ __forceinline__ double iexa_signp(double a) { return a >= 0.0 ? 1.0 : -1.0; }
// element address from a 32-bit index: ONE IMAD.WIDE.U32 (left to itself, the compiler widens the index arithmetic to
// 64 bits and spends 4-6 integer instructions per memory operation)
__device__ __forceinline__ const double* iexa_at(const double* p, unsigned i) {
  unsigned long long a;
  asm("mad.wide.u32 %0, %1, 8, %2;" : "=l"(a) : "r"(i), "l"((unsigned long long)p));
  return (const double*)a;
}
__device__ __forceinline__ double* iexa_atw(double* p, unsigned i) {
  unsigned long long a;
  asm("mad.wide.u32 %0, %1, 8, %2;" : "=l"(a) : "r"(i), "l"((unsigned long long)p));
  return (double*)a;
}
__device__ __forceinline__ double iexa_warp_sum(double s) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  return s;
}
)CUDA";

static std::string lit(double v) {
  char b[64];
  if (v != v) return "__longlong_as_double(0x7ff8000000000000LL)";
  if (v == INFINITY) return "__longlong_as_double(0x7ff0000000000000LL)";
  if (v == -INFINITY) return "__longlong_as_double(0xfff0000000000000LL)";
  snprintf(b, sizeof b, "%.17g", v);
  std::string s(b);
  if (s.find_first_of(".eEn") == std::string::npos) s += ".0";
  return s;
}

static const char *fn_name(int op) {
  switch (op) {
    case D_SQRT: return "sqrt"; case D_CBRT: return "cbrt"; case D_ABS: return "fabs";
    case D_SIGNP: return "iexa_signp"; case D_EXP: return "exp"; case D_EXP2: return "exp2";
    case D_LOG: return "log"; case D_LOG2: return "log2"; case D_LOG10: return "log10";
    case D_LOG1P: return "log1p"; case D_SIN: return "sin"; case D_COS: return "cos";
    case D_TAN: return "tan"; case D_ASIN: return "asin"; case D_ACOS: return "acos";
    case D_ATAN: return "atan"; case D_SINH: return "sinh"; case D_COSH: return "cosh";
    case D_TANH: return "tanh"; case D_ATANH: return "atanh";
    default: return nullptr;
  }
}

enum { SINK_DENSE = 0, SINK_SUM = 1, SINK_SCATTER = 2 };
constexpr int kMaxStagedStep = 64; // members with more slots than this store directly
// measured on B200 (config 3): capping registers at 64 (8 blocks of 128 threads per SM) lifts hess from 64% to 82% of the
// HBM roofline and cons/jac by 3 points; the occasional spill stays in L1
// products (round 2, config 3): hprod! with its rider body needs registers — 8 blocks/SM (64 registers, 400 B of spills) 0.291 ms,
// 6 blocks (80 registers, 336 B of stack) 0.215 ms, 4 blocks (128 registers, 88 B) 0.193 ms, 3 blocks (158 registers: its natural need,
// no spills) 0.236 ms — bandwidth needs the warps more than the last registers; jtprod! 0.310 -> 0.300 ms at 6; jprod 0.139 at 8 (0.144 at 10)
static const int kDefaultMinBlocks[KS__N] = {8, 8, 8, 8, 8, 8, 6, 6, 4, 6, 6};
int spec_block() {
  // threads per block of the specialised kernels (a multiple of 32).  Measured on B200, config 3:
  // 128 -> 0.585 ms/eval, 64 -> 0.597, 32 -> 0.627; on a 1/8 shard 64 and 128 tie (0.092 ms).
  static int b = [] { const char *e = getenv("IEXA_BLOCK"); int v = e ? atoi(e) : 128; return (v >= 32 && v <= 1024 && v % 32 == 0) ? v : 128; }();
  return b;
}
// IEXA_STAGE=tma: the staged warp tiles are written with cp.async.bulk (TMA, SASS UBLKCP) from two alternating
// shared-memory buffers instead of LDS+STG loops
static bool tma_stage() { const char *e = getenv("IEXA_STAGE"); return e && e[0] == 't'; }
// IEXA_STAGE=vec2: the staged warp tile is copied out with 128-bit LDS / STG (two doubles per lane and instruction).  The
// tile is shifted by one element in shared memory when its global start is 8-byte-odd, so that the 16-byte aligned pairs
// of the COO array are 16-byte aligned in shared memory too; an odd head / tail element is a plain store.  The
// load/store unit is the busiest unit of jac (l1tex 72 %): 35 STS + 35 LDS + 35 STG per thread become 35 + 18 + 18.
static bool vec2_stage() { const char *e = getenv("IEXA_STAGE"); return e && e[0] == 'v'; }
static bool block_stage() { const char *e = getenv("IEXA_STAGE"); return e && e[0] == 'b'; }
// Programmatic dependent launch (IEXA_PDL=0 disables).  Every specialised kernel signals
// griddepcontrol.launch_dependents as soon as it starts and executes griddepcontrol.wait before its first
// data access, and is launched with the programmatic-stream-serialisation attribute: the blocks of the NEXT
// callback become resident while the last wave of the previous kernel drains and start the moment it has
// completed and flushed.  The wait is unconditional, so ordering against whatever preceded the call on the
// caller's stream (the solver's own kernels writing x) is exactly that of a plain launch.
static bool use_pdl() { static bool v = [] { const char *e = getenv("IEXA_PDL"); return !(e && e[0] == '0'); }(); return v; }
// Loads first.  The fp64 sin / cos / tan routines contain branches (slow-path range reduction) and ptxas does not move
// loads across them: emitted in DAG order, a body's x / theta / column / y loads end up in 5-6 separate phases, each
// exposing a full memory latency per warp.  The first N loads of a body are therefore issued together at its top.
// N per callback, IEXA_HOIST="<obj>,<grad>,<cons>,<jac>,<hess>" (-1 = all).  Measured on B200, config 3 (ms):
//   hess 0.217 (N=0) -> 0.202 (8) -> 0.188 (12, all);  jac under a 48-register cap 0.224 (0) -> 0.230 (8), at 64
//   registers (8 blocks/SM) 0.2235 (0) -> 0.2222 (all);
//   cons under a 48-register cap (10 blocks/SM) 0.152 (0) -> 0.147 (8) -> 0.271 (all: 8 of the 19 values spill); with
//   the final kernels 0.134 (4) / 0.129 (8) / 0.132 (12) at 48 registers and 0.126 (all) at 64 registers (8 blocks/SM).
// slots 5..9 (jprod, jtprod phases 0/1, hprod phases 0/1): IEXA_HOIST_PROD="<jprod>,<jtprod>,<hprod>"
static int hoist_loads(int cb) {
  // products (B200, config 3, round 2): jprod 0.139 ms at 12 (0.148 at 24, 0.156 at 0); jtprod 0.310 at 24 (0.316 at 12, 0.344 at 0);
  // hprod (80 registers, rider body fused): 0.215 at 16
  static int v[KS__N] = {-1, -1, -1, -1, 16, 12, 24, 24, 16, 16, 16};
  static bool init = [] {
    if (const char *e = getenv("IEXA_HOIST")) {
      int w[5] = {v[0], v[1], v[2], v[3], v[4]};
      int n = sscanf(e, "%d,%d,%d,%d,%d", &w[0], &w[1], &w[2], &w[3], &w[4]);
      if (n == 1) for (int i = 1; i < 5; ++i) w[i] = w[0];
      for (int i = 0; i < 5; ++i) v[i] = w[i];
    }
    if (const char *e = getenv("IEXA_HOIST_PROD")) {
      int w[3] = {v[5], v[6], v[8]};
      sscanf(e, "%d,%d,%d", &w[0], &w[1], &w[2]);
      v[5] = w[0]; v[6] = v[7] = w[1]; v[8] = v[9] = w[2];
    }
    return true;
  }();
  (void)init;
  return v[cb] < 0 ? 1 << 30 : v[cb];
}
// 32-bit index arithmetic + mad.wide.u32 addressing per callback, IEXA_IDX32="<obj>,<grad>,<cons>,<jac>,<hess>".
// Measured on B200, config 3: cons 0.138 -> 0.130 ms (its collocation rows are 38 loads of pure index arithmetic),
// jac 0.226 -> 0.234 ms, hess unchanged: on for the value kernels only.
static bool idx32_for(int cb) {
  static int v[KS__N] = {1, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0};
  static bool init = [] {
    if (const char *e = getenv("IEXA_IDX32")) sscanf(e, "%d,%d,%d,%d,%d", &v[0], &v[1], &v[2], &v[3], &v[4]);
    return true;
  }();
  (void)init;
  return v[cb] != 0;
}
static bool prefetch_rest() { static bool v = [] { const char *e = getenv("IEXA_PREFETCH"); return e && e[0] == '1'; }(); return v; }
static bool pad_tiles() { const char *e = getenv("IEXA_PAD_TILES"); return e && e[0] == '1'; }
// Warp tiles: lane l writes its `step` slot values at wsm[l * stride + c].  With an EVEN step the 16 lanes of a half-warp
// hit only 8 / 4 / 2 of the 16 64-bit banks (step 2, 6: 2-way, step 4: 4-way conflicts — ncu counted 5.9 M conflicts on
// 4.7 M shared stores in jac of config 3, top stall mio_throttle); stride = step | 1 makes the writes conflict-free, the
// copy-out reads element j at j + j / step.  Measured on B200 (round 2): time-neutral on config 3 (jac 0.2233 padded vs
// 0.2225 ms) and slightly slower on the 118-bus OPF (hess 0.1278 vs 0.1248 ms) — the conflicts are real but the shared
// memory pipe is not what bounds these kernels; the extra index arithmetic of the copy-out costs as much as the
// conflicts did.  Off by default; IEXA_PAD_WARP_TILES=1 enables it.
static bool pad_warp_tiles() { static bool v = [] { const char *e = getenv("IEXA_PAD_WARP_TILES"); return e && e[0] == '1'; }(); return v; }

// Instances of a shape class handled by ONE block, one after the other (IEXA_CLASS_CHUNK, default 8).  A block that
// evaluates a single small generator at 128 supports lives for three dependent memory latencies (work item -> instance
// table -> x) and moves a few KB: thousands of such blocks are bound by their own lifetime (118-bus OPF x 10^4 scenarios:
// 224 000 blocks per callback).  Looping over a chunk of instances pays the block start once per chunk.
int class_chunk() { static int v = [] { const char *e = getenv("IEXA_CLASS_CHUNK"); int c = e ? atoi(e) : 8; return c >= 1 && c <= 1024 ? c : 8; }(); return v; }
// Shape-class bodies read, per instance, its index bases / output offsets (IT) and its literal constants (CT).  Staged once
// per block — one coalesced load of the whole chunk's tables into shared memory — they cost an LDS each instead of a
// uniform global load in front of every x load (work item -> instance table -> x was three dependent memory latencies
// per instance; now the table latency is paid once per chunk).  IEXA_CLASS_SMEM=0 keeps the global loads.
static bool cls_smem() { static bool v = [] { const char *e = getenv("IEXA_CLASS_SMEM"); return !(e && e[0] == '0'); }(); return v; }
static bool closed_form() { static bool v = [] { const char *e = getenv("IEXA_SCHED"); return !(e && e[0] == 't'); }(); return v; }

static CbSchedule make_schedule_impl(const Plan &plan, const std::vector<int> &groups, bool allow_big) {
  CbSchedule S;
  const int64_t B = spec_block();
  if (allow_big)
    for (int gi : groups) {
      const Group &G = plan.groups[gi];
      if (!G.is_class) S.NB = std::max<int64_t>(S.NB, (G.k1 - G.k0 + B - 1) / B);
    }
  if (S.NB > 60000) S.NB = 0; // gridDim.y limit: very long groups stay on the work table
  for (int gi : groups) {
    const Group &G = plan.groups[gi];
    const int64_t n = G.k1 - G.k0;
    int64_t sb = S.NB > 0 ? ((n + S.NB - 1) / S.NB + 31) / 32 * 32 : 0;
    // a cut that leaves less than a warp of supports per block (or nothing at all) is not worth NB blocks
    if (S.NB >= 64 && !G.is_class && n >= 16 * S.NB && sb <= B) { S.big.push_back(gi); S.sb.push_back((int)sb); }
    else {
      S.small.push_back(gi);
      S.ntable += ((n + B - 1) / B) * (G.is_class ? ((int64_t)G.n_inst() + class_chunk() - 1) / class_chunk() : 1);
    }
  }
  if (S.big.empty()) S.NB = 0;
  // 2-D grid: x = slot of a big group (fastest in dispatch order), y = block; table items fill the rows after NB
  S.gx = S.big.empty() ? std::max<int64_t>(1, std::min<int64_t>(S.ntable, 1024)) : (int64_t)S.big.size();
  S.gy = S.NB + (S.ntable + S.gx - 1) / S.gx;
  return S;
}

CbSchedule make_schedule(const Plan &plan, const std::vector<int> &groups) {
  CbSchedule S = make_schedule_impl(plan, groups, closed_form());
  // a few big groups next to a very long work table (thousands of shape-class instances) would need more grid rows than
  // gridDim.y allows: everything goes on the table then, which is laid out 1024 items per row
  if (S.gy > 65000 && !S.big.empty()) S = make_schedule_impl(plan, groups, false);
  return S;
}

// one group's case body for one callback
struct BodyGen {
  const Plan &P;
  const Group &G;
  int gid, prog, sink, cb = 0;
  std::vector<CiEntry> &ci;
  std::ostringstream o;
  std::map<int, std::string> colj;
  std::vector<std::string> ixname;
  bool use32, idx32;
  size_t smem_doubles = 0;
  bool rider = false;           // body appended to another group's case: k, kl, active, tid, ... are already in scope
  std::string outname = "out";  // objective groups inside the fused eval3 kernel write their Hessian slots to out3
  // fused eval3 body: outputs of member m and program p (0 value, 1 first, 2 second order) form the "virtual member" 3m+p
  bool fused3() const { return prog == GPROG_ALL; }
  int mem_of(int vm) const { return fused3() ? vm / 3 : vm; }
  int prog_of(int vm) const { return fused3() ? vm % 3 : prog; }
  std::string out_of(int vm) const { return fused3() ? (vm % 3 == 0 ? "out" : vm % 3 == 1 ? "out2" : "out3") : outname; }

  BodyGen(const Plan &P_, const Group &G_, int gid_, int prog_, int sink_, std::vector<CiEntry> &ci_)
      : P(P_), G(G_), gid(gid_), prog(prog_), sink(sink_), ci(ci_) {
    ixname.assign(G.ctx.uidx.size(), "");
    use32 = G.K < (1ll << 31);
    const int64_t lim = 1ll << 31;
    idx32 = use32 && P.nvar < lim && P.npar < lim && P.loc_ncon < lim && P.loc_nnzj < lim && P.loc_nnzh < lim;
  }
  std::string C(int kind, int a = 0, int b = 0) {
    if (G.is_class) { // per-instance table: [index bases..., then per member: row_local, out_val, out_d1, out_d2]
      const int nis = (int)G.ctx.uidx.size();
      // (IT / CT point into the chunk's tables staged in shared memory, or into global memory with IEXA_CLASS_SMEM=0)
      const std::string L = cls_smem() ? "IT[" : "__ldg(IT + ", R = cls_smem() ? "]" : ")";
      if (kind == CI_IDX_BASE) return L + std::to_string(a) + R;
      if (kind == CI_MEM_ROWLOC) return L + std::to_string(nis + 4 * a) + R;
      if (kind == CI_MEM_OUT) return L + std::to_string(nis + 4 * a + 1 + (b == PROG_D1 ? 1 : b == PROG_D2 ? 2 : 0)) + R;
    }
    for (size_t i = 0; i < ci.size(); ++i)
      if (ci[i].kind == kind && ci[i].group == gid && ci[i].a == a && ci[i].b == b) return "CI[" + std::to_string(i) + "]";
    ci.push_back(CiEntry{kind, gid, a, b});
    return "CI[" + std::to_string(ci.size() - 1) + "]";
  }
  std::string jvar(bool is_int, int slot) {
    int key = (is_int ? 0 : 100000) + slot;
    auto it = colj.find(key);
    if (it != colj.end()) return it->second;
    const Iterator &itr = P.itrs[G.itr];
    const ColRef &r = is_int ? itr.int_cols[G.ctx.int_cols[slot]] : itr.fp_cols[G.ctx.fp_cols[slot]];
    std::string name = std::string(is_int ? "ji" : "jf") + std::to_string(slot);
    const char *ty = use32 ? "unsigned" : "long long";
    std::string e = "kk";
    if (r.div != 1) e = "(" + e + " / " + std::to_string(r.div) + (use32 ? "u" : "ll") + ")";
    if (r.div * r.mod < G.K) e = "(" + e + " % " + std::to_string(r.mod) + (use32 ? "u" : "ll") + ")";
    o << "      const " << ty << " " << name << " = " << e << ";\n";
    colj[key] = name;
    return name;
  }
  std::string ix(int islot) {
    if (!ixname[islot].empty()) return ixname[islot];
    const IndexExpr &e = G.ctx.uidx[islot];
    std::string name = "ix" + std::to_string(islot);
    std::ostringstream ex;
    const Iterator &itr = P.itrs[G.itr];
    if (idx32) {
      // 32-bit index arithmetic (every index and offset of this plan fits): an address is then ONE integer add plus one
      // IMAD.WIDE, where 64-bit index arithmetic cost IADD3 + IMAD.X + LEA + LEA.HI.X per load (ncu source view of the
      // collocation rows: 195 integer instructions for 38 loads).  Unsigned wrap-around makes negative coefficients exact.
      ex << "(unsigned)" << C(CI_IDX_BASE, islot);
      for (auto &t : e.terms) {
        const ColRef &r = itr.int_cols[G.ctx.int_cols[t.first]];
        std::string j = "(unsigned)" + jvar(true, t.first);
        std::string val;
        const HostColumn &hc = P.columns[r.col];
        auto U = [](int64_t v) { return "(unsigned)(" + std::to_string(v) + "ll)"; };
        if (hc.iota) val = "(" + j + " + 1u)";
        else if (hc.affine) {
          val = "(" + U(hc.aa);
          if (hc.ab != 0) val += " + " + U(hc.ab) + " * (" + (hc.ac == 1 ? j : j + " / " + std::to_string(hc.ac) + "u") + ")";
          if (hc.ad != 0 && hc.ac > 1) val += " + " + U(hc.ad) + " * (" + j + " % " + std::to_string(hc.ac) + "u)";
          val += ")";
        } else val = "(unsigned)__ldg((const int*)" + C(CI_ICOL_PTR, t.first) + " + " + jvar(true, t.first) + ")";
        ex << " + " << U(t.second) << " * " << val;
      }
      o << "      const unsigned " << name << " = " << ex.str() << ";\n";
      ixname[islot] = name;
      return name;
    }
    ex << C(CI_IDX_BASE, islot);
    for (auto &t : e.terms) {
      const ColRef &r = itr.int_cols[G.ctx.int_cols[t.first]];
      std::string j = jvar(true, t.first);
      std::string val;
      const HostColumn &hc = P.columns[r.col];
      if (hc.iota) val = "((long long)" + j + " + 1)";
      else if (hc.affine) { // v(j) = aa + ab*(j / ac) + ad*(j % ac), literals: integer arithmetic instead of a column load
        const std::string u = use32 ? "u" : "ll";
        val = "(" + std::to_string(hc.aa) + "ll";
        if (hc.ab != 0) val += " + " + std::to_string(hc.ab) + "ll * (long long)(" + (hc.ac == 1 ? j : j + " / " + std::to_string(hc.ac) + u) + ")";
        if (hc.ad != 0 && hc.ac > 1) val += " + " + std::to_string(hc.ad) + "ll * (long long)(" + j + " % " + std::to_string(hc.ac) + u + ")";
        val += ")";
      }
      else val = "(long long)__ldg((const int*)" + C(CI_ICOL_PTR, t.first) + " + " + j + ")";
      ex << " + " << t.second << "ll * " << val;
    }
    o << "      const long long " << name << " = " << ex.str() << ";\n";
    ixname[islot] = name;
    return name;
  }
  std::string at(const std::string &arr, const std::string &i) { return idx32 ? "iexa_at(" + arr + ", " + i + ")" : "(" + arr + " + (" + i + "))"; }
  std::string atw(const std::string &arr, const std::string &i) { return idx32 ? "iexa_atw(" + arr + ", " + i + ")" : "(" + arr + " + (" + i + "))"; }
  // element offset  <constant-bank offset> + <per-thread part>  of a y load / direct store
  std::string off(const std::string &ci, const std::string &dyn) {
    return idx32 ? "(unsigned)" + ci + " + (unsigned)(" + dyn + ")" : ci + " + " + dyn;
  }

  std::string run() {
    const Program &pr = G.prog[prog];
    const auto &outmap = G.outmap[prog];
    // shared-memory staging of the members whose outputs are AoS tiles (ostep > 1).
    //   warp mode (default): every warp owns a 32 x maxstep scratch tile; a member's tile is written
    //   out by its own warp as soon as the member's last value exists (__syncwarp only, no block
    //   barrier, 32*maxstep*8 bytes per warp => shared memory never limits occupancy);
    //   block mode (IEXA_STAGE=block): all members' 128 x step tiles, one __syncthreads, copy at the end.
    const bool warp_mode = !block_stage();
    const size_t nvm = fused3() ? 3 * G.members.size() : G.members.size();
    std::vector<long long> soff(nvm, -1);
    std::vector<int> stride(nvm, 0), step(nvm, 1), outs_left(nvm, 0);
    int maxstep = 0;
    for (size_t m = 0; m < nvm; ++m) {
      const Generator &g = P.member(G, mem_of((int)m));
      step[m] = prog_of((int)m) == PROG_D1 ? g.c.o1step : prog_of((int)m) == PROG_D2 ? g.c.o2step : 1;
      if (sink == SINK_DENSE && step[m] > 1 && step[m] <= kMaxStagedStep) {
        if (warp_mode) {
          stride[m] = (pad_warp_tiles() && !tma_stage() && !vec2_stage()) ? (step[m] | 1) : step[m];
          soff[m] = 0;
          maxstep = std::max(maxstep, stride[m]);
        }
        else {
          stride[m] = pad_tiles() ? (step[m] | 1) : step[m];
          soff[m] = (long long)smem_doubles;
          smem_doubles += (size_t)stride[m] * spec_block();
        }
      }
    }
    const bool tma = warp_mode && tma_stage();
    const bool vec2 = warp_mode && !tma && vec2_stage();
    const int bufsz = 32 * maxstep + 2; // per warp buffer (+2: the tile is shifted by one element when its global start is 8-byte-odd)
    if (warp_mode) smem_doubles = tma ? (size_t)2 * bufsz * (spec_block() / 32) : vec2 ? (size_t)bufsz * (spec_block() / 32) : (size_t)maxstep * 32 * (spec_block() / 32);
    int nflush = 0;
    for (auto &om : outmap) outs_left[om.first]++;
    std::vector<std::vector<std::pair<int, std::string>>> pending(nvm);
    if (!rider) {
    if (G.is_class) {
      o << "      const int nblk_ = (int)" << C(CI_CLS_NBLK) << ";\n"
        << "      const int chunk_ = blk_ / nblk_, wb_ = blk_ - chunk_ * nblk_;\n";
    } else {
      o << "      const int wb_ = blk_;\n";
    }
    o << "      const long long k0 = " << C(CI_K0) << ", k1 = " << C(CI_K1) << ";\n"
      << "      const int SB_ = (int)" << C(CI_SB, cb) << ";\n"
      << "      const long long kb = k0 + (long long)wb_ * SB_;\n"
      << "      if (kb >= k1) break;\n"   // closed-form schedule: trailing blocks of a group may be empty
      << "      const int nact = (int)((k1 - kb) < SB_ ? (k1 - kb) : SB_);\n"
      << "      const bool active = tid < nact;\n"
      << "      const long long k = kb + (active ? tid : nact - 1);\n"
      << "      const long long kl = k - k0; (void)kl;\n";
    }
    if (tma && maxstep > 0)
      o << "      double* __restrict__ wsm = sm + (tid >> 5) * " << 2 * bufsz << ";\n"
        << "      const int out_par = (int)(((unsigned long long)out >> 3) & 1ull);\n"
        << "      const int lane = tid & 31, w0 = tid & ~31;\n"
        << "      const int wact = nact - w0 < 0 ? 0 : (nact - w0 > 32 ? 32 : nact - w0);\n";
    else if (vec2 && maxstep > 0)
      o << "      double* __restrict__ wsm = sm + (tid >> 5) * " << bufsz << ";\n"
        << "      const int out_par = (int)(((unsigned long long)out >> 3) & 1ull);\n"
        << "      const int lane = tid & 31, w0 = tid & ~31;\n"
        << "      const int wact = nact - w0 < 0 ? 0 : (nact - w0 > 32 ? 32 : nact - w0);\n";
    else if (warp_mode && maxstep > 0)
      o << "      double* __restrict__ wsm = sm + (tid >> 5) * " << 32 * maxstep << ";\n"
        << "      const int lane = tid & 31, w0 = tid & ~31;\n"
        << "      const int wact = nact - w0 < 0 ? 0 : (nact - w0 > 32 ? 32 : nact - w0);\n";
    if (!rider) {
      if (use32) o << "      const unsigned kk = (unsigned)k; (void)kk;\n";
      else o << "      const long long kk = k; (void)kk;\n";
    }
    if (G.is_class) { // the instances of this block's chunk, one after the other (same supports, same program)
      const int ninst = (int)G.n_inst(), ch = class_chunk();
      const char *ue = getenv("IEXA_CLASS_UNROLL");
      // two instances in flight (IEXA_CLASS_UNROLL): the loads of the next instance overlap the arithmetic and the
      // tile write-out of this one — 118-bus OPF: 0.312 (1) -> 0.290 (2) -> 0.288 ms (4)
      const size_t W = G.ctx.uidx.size() + 4 * G.members.size(), NC = std::max<size_t>(G.cpar_nodes.size(), 1);
      if (cls_smem()) {
        const size_t off = smem_doubles;
        smem_doubles += (size_t)ch * (W + NC);
        o << "      long long* __restrict__ sIT_ = (long long*)(sm + " << off << ");\n"
          << "      double* __restrict__ sCT_ = sm + " << off + (size_t)ch * W << ";\n"
          << "      { const int i0_ = chunk_ * " << ch << ", n_ = (" << ninst << " - i0_) < " << ch << " ? (" << ninst << " - i0_) : " << ch << ";\n"
          << "        const long long* __restrict__ gI_ = (const long long*)" << C(CI_CLS_ITAB) << " + (long long)i0_ * " << W << ";\n"
          << "        const double* __restrict__ gC_ = (const double*)" << C(CI_CLS_DTAB) << " + (long long)i0_ * " << NC << ";\n"
          << "        for (int i_ = tid; i_ < n_ * " << W << "; i_ += IEXA_BLOCK) sIT_[i_] = __ldg(gI_ + i_);\n"
          << "        for (int i_ = tid; i_ < n_ * " << NC << "; i_ += IEXA_BLOCK) sCT_[i_] = __ldg(gC_ + i_);\n"
          << "        __syncthreads(); }\n";
      }
      o << "      _Pragma(\"unroll " << (ue ? atoi(ue) : 2) << "\")\n"
        << "      for (int inst_ = chunk_ * " << ch << "; inst_ < " << ninst << " && inst_ < (chunk_ + 1) * " << ch << "; ++inst_) {\n";
      if (cls_smem())
        o << "      const long long* __restrict__ IT = sIT_ + (inst_ - chunk_ * " << ch << ") * " << W << ";\n"
          << "      const double* __restrict__ CT = sCT_ + (inst_ - chunk_ * " << ch << ") * " << NC << "; (void)CT; (void)IT;\n";
      else
        o << "      const long long* __restrict__ IT = (const long long*)" << C(CI_CLS_ITAB) << " + (long long)inst_ * " << W << ";\n"
          << "      const double* __restrict__ CT = (const double*)" << C(CI_CLS_DTAB) << " + (long long)inst_ * " << NC << "; (void)CT; (void)IT;\n";
    }

    std::vector<std::string> cur(pr.nreg > 0 ? pr.nreg : 1);
    auto opd = [&](int a) -> std::string { return a >= 0 ? cur[a] : lit(pr.cpool[~a]); };
    // a load has no register operands: its line can be emitted anywhere before its first use
    auto emit_load = [&](const Instr &I, const std::string &t) {
      switch (I.op) {
        case D_FIELD: {
          std::string j = jvar(false, I.a);
          o << "      const double " << t << " = __ldg((const double*)" << C(CI_FCOL_PTR, I.a) << " + " << j << ");\n";
          break;
        }
        case D_LOADX: { std::string i = ix(I.a); o << "      const double " << t << " = __ldg(" << at("x", i + " - 1") << ");\n"; break; }
        case D_LOADP: { std::string i = ix(I.a); o << "      const double " << t << " = __ldg(" << at("theta", i + " - 1") << ");\n"; break; }
        case D_LOADV: { std::string i = ix(I.a); o << "      const double " << t << " = __ldg(" << at("v", i + " - 1") << ");\n"; break; }
        case D_W:
          if (G.is_obj) o << "      const double " << t << " = sigma;\n";
          else o << "      const double " << t << " = y ? __ldg(" << at("y", off(C(CI_MEM_ROWLOC, I.a), "kl")) << ") : 0.0;\n";
          break;
        case D_CPAR: o << "      const double " << t << " = " << (cls_smem() ? "CT[" + std::to_string(I.a) + "]" : "__ldg(CT + " + std::to_string(I.a) + ")") << ";\n"; break;
        default: break;
      }
    };
    auto is_load = [](int op) { return op == D_FIELD || op == D_LOADX || op == D_LOADP || op == D_LOADV || op == D_W || op == D_CPAR; };
    std::vector<char> hoisted(pr.code.size(), 0);
    {
      int nh = 0, q = 0;
      for (const Instr &I : pr.code) {
        if (is_load(I.op) && nh < hoist_loads(cb)) { emit_load(I, "t" + std::to_string(q)); hoisted[q] = 1; ++nh; }
        ++q;
      }
      // IEXA_PREFETCH=1: loads beyond the hoisting budget are announced to the L2 at the top of the body
      // (prefetch.global.L2 has no destination register)
      if (prefetch_rest()) {
        q = 0;
        for (const Instr &I : pr.code) {
          if (!hoisted[q]) {
            std::string ptr;
            if (I.op == D_LOADX) ptr = at("x", ix(I.a) + " - 1");
            else if (I.op == D_LOADV) ptr = at("v", ix(I.a) + " - 1");
            else if (I.op == D_LOADP) ptr = at("theta", ix(I.a) + " - 1");
            else if (I.op == D_FIELD) ptr = "((const double*)" + C(CI_FCOL_PTR, I.a) + " + " + jvar(false, I.a) + ")";
            else if (I.op == D_W && !G.is_obj) ptr = at("y", off(C(CI_MEM_ROWLOC, I.a), "kl"));
            if (!ptr.empty()) o << "      asm volatile(\"prefetch.global.L2 [%0];\" :: \"l\"(" << ptr << "));\n";
          }
          ++q;
        }
      }
    }
    // sin and cos of the SAME value (every rotation angle of the quadrotor, every voltage angle difference of the OPF
    // rows) become one sincos(): one range reduction, one slow-path branch.  partner[pc] = pc of the other half.
    std::vector<int> partner(pr.code.size(), -1);
    if (!getenv("IEXA_NO_SINCOS")) {
      std::vector<std::string> nm(pr.nreg > 0 ? pr.nreg : 1);
      std::map<std::string, std::pair<int, int>> sc; // operand name -> (pc of sin, pc of cos)
      int q = 0;
      for (const Instr &I : pr.code) {
        if (I.op == D_SIN || I.op == D_COS) {
          const std::string a = I.a >= 0 ? nm[I.a] : lit(pr.cpool[~I.a]);
          auto it = sc.find(a);
          if (it == sc.end()) it = sc.emplace(a, std::make_pair(-1, -1)).first;
          int &slot = I.op == D_SIN ? it->second.first : it->second.second;
          if (slot < 0) slot = q;
        }
        if (I.op != D_OUT) nm[I.dst] = "t" + std::to_string(q);
        ++q;
      }
      for (auto &kv : sc)
        if (kv.second.first >= 0 && kv.second.second >= 0) { partner[kv.second.first] = kv.second.second; partner[kv.second.second] = kv.second.first; }
    }
    int pc = 0;
    for (const Instr &I : pr.code) {
      std::string t = "t" + std::to_string(pc++);
      if ((I.op == D_SIN || I.op == D_COS) && partner[pc - 1] >= 0) {
        const int other = partner[pc - 1];
        if (other > pc - 1) { // first of the pair: define both names
          const std::string to = "t" + std::to_string(other);
          o << "      double " << t << ", " << to << "; sincos(" << opd(I.a) << ", &" << (I.op == D_SIN ? t : to) << ", &" << (I.op == D_SIN ? to : t) << ");\n";
        }
        cur[I.dst] = t;
        continue;
      }
      switch (I.op) {
        case D_FIELD: case D_LOADX: case D_LOADP: case D_LOADV: case D_W: case D_CPAR:
          if (!hoisted[pc - 1]) emit_load(I, t);
          cur[I.dst] = t; break;
        case D_SELNE: { std::string a = ix(I.a), b = ix(I.b); o << "      const double " << t << " = (" << a << " != " << b << ") ? 1.0 : 0.0;\n"; cur[I.dst] = t; break; }
        case D_SEL2: { std::string a = ix(I.a), b = ix(I.b); o << "      const double " << t << " = (" << a << " == " << b << ") ? 2.0 : 1.0;\n"; cur[I.dst] = t; break; }
        case D_OUT: {
          std::string v = opd(I.a);
          int m = outmap[I.dst].first, c = outmap[I.dst].second;
          if (sink == SINK_DENSE) {
            if (soff[m] >= 0 && warp_mode) pending[m].push_back({c, v}); // flushed when the member is complete
            else if (soff[m] >= 0) o << "      sm[" << soff[m] << " + tid * " << stride[m] << " + " << c << "] = " << v << ";\n";
            else if (step[m] == 1) o << "      if (active) __stcs(" << atw(out_of(m), off(C(CI_MEM_OUT, mem_of(m), prog_of(m)), "kl")) << ", " << v << ");\n";
            else o << "      if (active) __stcs(" << atw(out_of(m), off(C(CI_MEM_OUT, mem_of(m), prog_of(m)), "kl * " + std::to_string(step[m]) + " + " + std::to_string(c))) << ", " << v << ");\n";
            if (--outs_left[m] == 0 && soff[m] >= 0 && tma) {
              // TMA write-out: the tile sits in buffer (flush & 1), shifted by `par` so that the 16-byte aligned
              // part of the global range is 16-byte aligned in shared memory too; lane 0 issues ONE bulk copy,
              // the odd head / tail elements (at most one each) are plain stores
              const int f = nflush++;
              if (f >= 2) o << "      if (lane == 0) asm volatile(\"cp.async.bulk.wait_group.read 1;\" ::: \"memory\");\n      __syncwarp();\n";
              o << "      { const long long e0 = " << C(CI_MEM_OUT, m, prog) << " + (kb - k0 + w0) * " << step[m] << ";\n"
                << "        const int par = (int)((e0 + out_par) & 1);\n"
                << "        double* __restrict__ buf = wsm + " << (f & 1) * bufsz << " + par;\n";
              for (auto &pv : pending[m]) o << "        buf[lane * " << step[m] << " + " << pv.first << "] = " << pv.second << ";\n";
              o << "        asm volatile(\"fence.proxy.async.shared::cta;\" ::: \"memory\");\n"
                << "        __syncwarp();\n"
                << "        if (lane == 0) {\n"
                << "          double* __restrict__ dst = out + e0;\n"
                << "          const int n = wact * " << step[m] << ";\n"
                << "          const int head = n > 0 ? par : 0, body = (n - head) & ~1, tail = n - head - body;\n"
                << "          if (head) dst[0] = buf[0];\n"
                << "          if (tail) dst[n - 1] = buf[n - 1];\n"
                << "          if (body > 0) asm volatile(\"cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\" :: \"l\"(dst + head), \"r\"((unsigned)__cvta_generic_to_shared(buf + head)), \"r\"(body * 8) : \"memory\");\n"
                << "          asm volatile(\"cp.async.bulk.commit_group;\" ::: \"memory\");\n"
                << "        }\n      }\n";
            } else if (outs_left[m] == 0 && soff[m] >= 0 && vec2) {
              o << "      { const long long e0 = " << C(CI_MEM_OUT, m, prog) << " + (kb - k0 + w0) * " << step[m] << ";\n"
                << "        const int par = (int)((e0 + out_par) & 1);\n"
                << "        double* __restrict__ buf = wsm + par;\n";
              for (auto &pv : pending[m]) o << "        buf[lane * " << step[m] << " + " << pv.first << "] = " << pv.second << ";\n";
              o << "        __syncwarp();\n"
                << "        double* __restrict__ dst = out + e0;\n"
                << "        const int n = wact * " << step[m] << ";\n"
                << "        const int head = n > 0 ? par : 0, npair = (n - head) >> 1;\n"
                << "        if (lane == 0 && head) dst[0] = buf[0];\n"
                << "        if (lane == 1 && ((n - head) & 1)) dst[n - 1] = buf[n - 1];\n"
                << "        const double2* __restrict__ s2 = reinterpret_cast<const double2*>(buf + head);\n"
                << "        double2* __restrict__ d2 = reinterpret_cast<double2*>(dst + head);\n"
                << "        _Pragma(\"unroll\")\n"
                << "        for (int i_ = 0; i_ < " << (step[m] + 1) / 2 << "; ++i_) { const int p_ = lane + 32 * i_; if (p_ < npair) d2[p_] = s2[p_]; }\n"
                << "        __syncwarp(); }\n";
            } else if (outs_left[m] == 0 && soff[m] >= 0 && warp_mode) {
              // the member's tile is complete in this warp: 32*step contiguous doubles of the COO array
              for (auto &pv : pending[m]) o << "      wsm[lane * " << stride[m] << " + " << pv.first << "] = " << pv.second << ";\n";
              o << "      __syncwarp();\n"
                << "      { double* __restrict__ dst = " << out_of(m) << " + (" << C(CI_MEM_OUT, mem_of(m), prog_of(m)) << " + (kb - k0 + w0) * " << step[m] << ");\n"
                << "        const int n = wact * " << step[m] << ";\n"
                << "        _Pragma(\"unroll\")\n"
                << "        for (int i_ = 0; i_ < " << step[m] << "; ++i_) { const int j = lane + 32 * i_; if (j < n) dst[j] = wsm["
                << (stride[m] == step[m] ? std::string("j") : "j + j / " + std::to_string(step[m])) << "]; } }\n"
                << "      __syncwarp();\n";
            }
          } else if (sink == SINK_SUM) {
            o << "      if (active) acc += " << v << ";\n";
          } else {
            int islot;
            bool direct = true; // single writer for every generator that runs this body
            if (prog == PROG_JTV || prog == PROG_HV) { // scatter products: (0, index slot); Plan::analyse_scatter
              islot = c;
              const int w = prog == PROG_JTV ? 0 : 1;
              direct = G.scat_phase[w] == 0 && (size_t)I.dst < G.scat_direct[w].size() && G.scat_direct[w][I.dst];
              if (direct && G.scat_direct[w][I.dst] == 2) { // rider: the same thread stored this entry a moment ago
                std::string i = ix(islot);
                o << "      if (active) out[" << i << " - 1] += " << v << ";\n";
                break;
              }
            } else {
              islot = G.jac_slot[m][c];
              const std::vector<Generator> &gens = G.is_obj ? P.objs : P.cons;
              std::vector<int32_t> who{G.members[m]};
              if (G.is_class) { who.clear(); for (size_t q = 0; q < G.n_inst(); ++q) who.push_back(G.inst_gen(q, m)); }
              for (int32_t gi : who) direct = direct && (size_t)c < gens[gi].grad_direct.size() && gens[gi].grad_direct[c];
            }
            if (direct) {
              std::string i = ix(islot);
              o << "      if (active) out[" << i << " - 1] = " << v << ";\n";
            } else if (G.ctx.uidx[islot].terms.empty()) {
              o << "      { const double s_ = iexa_warp_sum(active ? " << v << " : 0.0); if ((tid & 31) == 0) atomicAdd(out + ("
                << C(CI_IDX_BASE, islot) << " - 1), s_); }\n";
            } else {
              std::string i = ix(islot);
              // supports 2e and 2e+1 address the SAME entry when every integer column of the index repeats each value
              // twice (the lower-bound index of the two rows of a collocation element, transform.jl:485-505): the two
              // lanes are summed with one shuffle and the even lane issues ONE atomic — jtprod! of config 3 is bound by
              // the L2's RED sector rate (ncu: 37.5 M RED sectors, issue slots 10 % busy), 27 of its 36 atomics per row pair up
              bool pairs = (prog == PROG_JTV || prog == PROG_HV) && !getenv("IEXA_NO_PAIR_REDUCE");
              for (auto &t : G.ctx.uidx[islot].terms) {
                const ColRef &r = P.itrs[G.itr].int_cols[G.ctx.int_cols[t.first]];
                const HostColumn &hc = P.columns[r.col];
                pairs = pairs && hc.affine && hc.ac == 2 && hc.ad == 0 && r.div == 1 && r.mod == G.K;
              }
              if (pairs)
                o << "      { const double s_ = active ? " << v << " : 0.0; const double o_ = __shfl_xor_sync(0xffffffffu, s_, 1);\n"
                  << "        if (!(kb & 1)) { if (!(tid & 1) && active) atomicAdd(out + (" << i << " - 1), s_ + o_); }\n"
                  << "        else if (active) atomicAdd(out + (" << i << " - 1), s_); }\n";
              else
                o << "      if (active) atomicAdd(out + (" << i << " - 1), " << v << ");\n";
            }
          }
          break;
        }
        case D_ADD: o << "      const double " << t << " = " << opd(I.a) << " + " << opd(I.b) << ";\n"; cur[I.dst] = t; break;
        case D_SUB: o << "      const double " << t << " = " << opd(I.a) << " - " << opd(I.b) << ";\n"; cur[I.dst] = t; break;
        case D_MUL: o << "      const double " << t << " = " << opd(I.a) << " * " << opd(I.b) << ";\n"; cur[I.dst] = t; break;
        case D_DIV: o << "      const double " << t << " = " << opd(I.a) << " / " << opd(I.b) << ";\n"; cur[I.dst] = t; break;
        case D_NEG: o << "      const double " << t << " = -" << opd(I.a) << ";\n"; cur[I.dst] = t; break;
        case D_POW: o << "      const double " << t << " = pow(" << opd(I.a) << ", " << opd(I.b) << ");\n"; cur[I.dst] = t; break;
        default: {
          const char *f = fn_name(I.op);
          o << "      const double " << t << " = " << (f ? f : "/*?*/") << "(" << opd(I.a) << ");\n";
          cur[I.dst] = t;
        }
      }
    }
    if (tma && nflush > 0) // shared memory must outlive the bulk reads
      o << "      if (lane == 0) asm volatile(\"cp.async.bulk.wait_group.read 0;\" ::: \"memory\");\n      __syncwarp();\n";
    // coalesced write-out of the staged tiles
    bool any = false;
    for (size_t m = 0; m < G.members.size(); ++m) any = any || soff[m] >= 0;
    if (any && !warp_mode) {
      o << "      __syncthreads();\n";
      for (size_t m = 0; m < G.members.size(); ++m) {
        if (soff[m] < 0) continue;
        o << "      { double* __restrict__ dst = out + (" << C(CI_MEM_OUT, (int)m, prog) << " + (kb - k0) * " << step[m] << ");\n"
          << "        const int n = nact * " << step[m] << ";\n"
          << "        _Pragma(\"unroll 4\")\n"
          << "        for (int j = tid; j < n; j += IEXA_BLOCK) ";
        if (stride[m] == step[m]) o << "dst[j] = sm[" << soff[m] << " + j];";
        else o << "{ const int t_ = j / " << step[m] << "; dst[j] = sm[" << soff[m] << " + j + t_]; }";
        o << " }\n";
      }
      if (G.is_class) o << "      __syncthreads();\n"; // the tiles are reused by the next instance of the chunk
    }
    if (G.is_class) o << "      }\n";
    return o.str();
  }
};

static const char *kCbName[KS__N] = {"iexa_cb_obj", "iexa_cb_grad", "iexa_cb_cons", "iexa_cb_jac", "iexa_cb_hess",
                                      "iexa_cb_jprod", "iexa_cb_jtprod_p0", "iexa_cb_jtprod_p1", "iexa_cb_hprod_p0", "iexa_cb_hprod_p1",
                                      "iexa_cb_eval3"};

GeneratedSource generate_source(const Plan &plan, int set) {
  GeneratedSource out;
  std::ostringstream kernels;
  struct CbDef { int ks, prog, sink; bool obj, con; int phase; };
  const CbDef defs[KS__N] = {{CB_OBJ, PROG_VAL, SINK_SUM, true, false, -1}, {CB_GRAD, PROG_D1, SINK_SCATTER, true, false, -1},
                             {CB_CONS, PROG_VAL, SINK_DENSE, false, true, -1}, {CB_JAC, PROG_D1, SINK_DENSE, false, true, -1},
                             {CB_HESS, PROG_D2, SINK_DENSE, true, true, -1},
                             {KS_JPROD, PROG_JV, SINK_DENSE, false, true, -1},
                             {KS_JTPROD0, PROG_JTV, SINK_SCATTER, false, true, 0}, {KS_JTPROD1, PROG_JTV, SINK_SCATTER, false, true, 1},
                             {KS_HPROD0, PROG_HV, SINK_SCATTER, true, true, 0}, {KS_HPROD1, PROG_HV, SINK_SCATTER, true, true, 1},
                             // cons! + jac_coord! + hess_coord! in one launch: constraint groups run their fused program
                             // (GPROG_ALL) into out / out2 / out3, objective groups their second-order program into out3
                             {KS_EVAL3, GPROG_ALL, SINK_DENSE, true, true, -1}};
  for (const CbDef &d : defs) {
    if (ks_set(d.ks) != set) continue;
    if (d.ks == KS_EVAL3 && (block_stage() || tma_stage() || vec2_stage())) continue; // default warp staging only
    std::ostringstream cases;
    size_t smem = 0;
    bool any = false;
    for (size_t gi = 0; gi < plan.groups.size(); ++gi) {
      const Group &G = plan.groups[gi];
      if (G.is_obj ? !d.obj : !d.con) continue;
      int gprog = d.prog;
      if (d.ks == KS_EVAL3 && G.is_obj) gprog = PROG_D2;
      if (gprog != PROG_VAL && G.prog[gprog].nout == 0) continue;
      if (d.phase >= 0 && G.scat_phase[d.prog == PROG_JTV ? 0 : 1] != d.phase) continue;
      const int sw = d.prog == PROG_JTV ? 0 : 1;
      if (d.phase == 0 && G.scat_rider_of[sw] >= 0) continue; // evaluated inside its primary's case (below)
      BodyGen bg(plan, G, (int)gi, gprog, d.sink, out.ci);
      if (d.ks == KS_EVAL3 && G.is_obj) bg.outname = "out3";
      bg.cb = d.ks;
      bg.idx32 = bg.idx32 && idx32_for(d.ks);
      std::string body = bg.run();
      if (d.phase == 0)
        for (size_t ri = 0; ri < plan.groups.size(); ++ri) {
          const Group &R = plan.groups[ri];
          if (R.scat_rider_of[sw] != (int)gi || R.scat_phase[sw] != 0 || R.prog[d.prog].nout == 0) continue;
          BodyGen rb(plan, R, (int)ri, d.prog, d.sink, out.ci);
          rb.cb = d.ks;
          rb.rider = true;
          rb.use32 = bg.use32;           // the primary declared kk with its own width
          rb.idx32 = bg.idx32;
          body += "      { // rider: group " + std::to_string(ri) + " on the same supports, same thread\n" + rb.run() + "      }\n";
        }
      cases << "    case " << gi << ":\n    {\n" << body << "    } break;\n";
      smem = std::max(smem, bg.smem_doubles * 8);
      out.groups_of[d.ks].push_back((int)gi);
      any = true;
    }
    if (!any) continue;
    out.smem_bytes[d.ks] = smem;
    out.sched[d.ks] = make_schedule(plan, out.groups_of[d.ks]);
    const size_t sbase = out.ci.size(); // [#closed-form blocks, #big groups, group id of slot 0, 1, ...]
    for (int j = 0; j < 2 + (int)out.groups_of[d.ks].size(); ++j) out.ci.push_back(CiEntry{CI_SCHED, -1, d.ks, j});
    // optional occupancy hint: IEXA_MINBLOCKS="<obj>,<grad>,<cons>,<jac>,<hess>" (0 = let ptxas decide)
    int minb = 0;
    const char *mbe = getenv("IEXA_MINBLOCKS");
    if (mbe && d.ks < 5) {
      int v[5] = {0, 0, 0, 0, 0};
      sscanf(mbe, "%d,%d,%d,%d,%d", &v[0], &v[1], &v[2], &v[3], &v[4]);
      minb = v[d.ks];
    } else {
      minb = kDefaultMinBlocks[d.ks] * 128 / spec_block();
      if (const char *pe = getenv("IEXA_MINBLOCKS_PROD")) { // "<jprod>,<jtprod>,<hprod>" (tuning)
        int v[3] = {8, 8, 8};
        sscanf(pe, "%d,%d,%d", &v[0], &v[1], &v[2]);
        if (d.ks == KS_JPROD) minb = v[0];
        else if (d.ks == KS_JTPROD0 || d.ks == KS_JTPROD1) minb = v[1];
        else if (d.ks == KS_HPROD0 || d.ks == KS_HPROD1) minb = v[2];
      }
    }
    kernels << "extern \"C\" __global__ void __launch_bounds__(IEXA_BLOCK" << (minb > 0 ? ", " + std::to_string(minb) : std::string()) << ") " << kCbName[d.ks]
            << "(const WorkItem* __restrict__ work, const double* __restrict__ x,\n"
               "    const double* __restrict__ theta, const double* __restrict__ y, const double* __restrict__ v, double sigma,\n"
               "    double* __restrict__ out, double* __restrict__ partials, double* __restrict__ out2, double* __restrict__ out3) {\n"
               "  extern __shared__ __align__(16) double sm[];\n"
               "  const int tid = threadIdx.x;\n"
               "  double acc = 0.0;\n"
            << (use_pdl() ? "  asm volatile(\"griddepcontrol.launch_dependents;\" ::: \"memory\");\n" : "")
            << "  // block -> (group, block of supports): blockIdx.y < NB: big group of slot blockIdx.x, block blockIdx.y;\n"
               "  // the rows after NB walk the work table of the small / shape-class groups\n"
               "  int gen_ = -1, blk_ = 0;\n"
               "  const unsigned lin_ = blockIdx.y * gridDim.x + blockIdx.x; (void)lin_;\n"
               "  { const unsigned nb_ = (unsigned)CI[" << sbase << "];\n"
               "    if (blockIdx.y < nb_) { blk_ = (int)blockIdx.y; gen_ = (int)CI[" << sbase + 2 << " + blockIdx.x]; }\n"
               "    else { const unsigned t_ = (blockIdx.y - nb_) * gridDim.x + blockIdx.x;\n"
               "           if (t_ < (unsigned)CI[" << sbase + 1 << "]) { const WorkItem w = work[t_]; gen_ = w.gen; blk_ = w.blk; } } }\n"
            << (use_pdl() ? "  asm volatile(\"griddepcontrol.wait;\" ::: \"memory\");\n" : "")
            <<
               "  (void)theta; (void)x; (void)y; (void)v; (void)sigma; (void)sm; (void)out2; (void)out3;\n"
               "  switch (gen_) {\n"
            << cases.str() << "    default: break;\n  }\n";
    if (d.sink == SINK_SUM) {
      kernels << "  __shared__ double red[IEXA_BLOCK / 32];\n"
                 "  acc = iexa_warp_sum(acc);\n"
                 "  if ((tid & 31) == 0) red[tid >> 5] = acc;\n"
                 "  __syncthreads();\n"
                 "  if (tid == 0) { double s = 0.0; for (int i = 0; i < IEXA_BLOCK / 32; ++i) s += red[i]; partials[lin_] = s; }\n";
    } else {
      kernels << "  (void)acc; (void)partials;\n";
    }
    kernels << "}\n\n";
  }
  std::ostringstream src;
  src << "#define IEXA_BLOCK " << spec_block() << "\n" << kPrelude << "__constant__ long long CI[" << std::max<size_t>(out.ci.size(), 1) << "];\n\n" << kernels.str();
  out.text = src.str();
  return out;
}

std::string Specialiser::generate_source(const Plan &plan, int set) { return iexa::generate_source(plan, set).text; }

// Generated sources depend on the STRUCTURE of a model only (every size, offset and pointer sits in the constant table),
// so a re-build of the same model with other supports / another shard — the closed-loop re-solves of test/solve.jl:134-209
// — produces the same text: compiled images are kept per process, keyed by the source.
static std::map<std::string, std::vector<char>> &cubin_cache() { static std::map<std::string, std::vector<char>> c; return c; }
static std::mutex &cubin_cache_mutex() { static std::mutex m; return m; }

// ---- persistent cubin cache -------------------------------------------------------------------------------------
// NVRTC costs 1-5 s per model (config 3: 1.2 s of a 1.7 s iexa_finalize; 118-bus OPF: 4 s) and the generated text depends
// only on the model's STRUCTURE, so the compiled image is also kept on disk, keyed by two independent 64-bit hashes of
// (source, compiler options, NVRTC version): a re-run of the same model — every solve of a parameter study, every rank
// of a multi-GPU job after the first — loads the image instead of compiling.  Directory: $IEXA_CACHE_DIR, else
// $XDG_CACHE_HOME/iexa_b200, else ~/.cache/iexa_b200, else /tmp/iexa_b200-cache; IEXA_CACHE_DIR=off disables it.
// Files are written to a temporary name and renamed (atomic for concurrent ranks); a file whose header does not match
// the source length and second hash is ignored.
static uint64_t fnv1a(const std::string &s, uint64_t h) {
  for (unsigned char c : s) { h ^= c; h *= 1099511628211ull; }
  return h;
}
static std::string cache_dir() {
  const char *e = getenv("IEXA_CACHE_DIR");
  if (e && (std::string(e) == "off" || std::string(e) == "0")) return "";
  std::string d;
  if (e && *e) d = e;
  else if (const char *x = getenv("XDG_CACHE_HOME")) d = std::string(x) + "/iexa_b200";
  else if (const char *h = getenv("HOME")) d = std::string(h) + "/.cache/iexa_b200";
  else d = "/tmp/iexa_b200-cache";
  std::string cmd; // mkdir -p without <filesystem> (nvcc host compilers differ): create each prefix
  for (size_t i = 1; i <= d.size(); ++i)
    if (i == d.size() || d[i] == '/') { cmd = d.substr(0, i); ::mkdir(cmd.c_str(), 0700); }
  return d;
}
struct CacheHeader { char magic[8]; uint64_t src_len, hash2, cubin_len; };
static bool disk_cache_load(const std::string &key_text, std::vector<char> &cubin) {
  const std::string dir = cache_dir();
  if (dir.empty()) return false;
  char name[64];
  snprintf(name, sizeof name, "/%016llx.cubin", (unsigned long long)fnv1a(key_text, 1469598103934665603ull));
  std::ifstream f(dir + name, std::ios::binary);
  if (!f) return false;
  CacheHeader h{};
  f.read((char *)&h, sizeof h);
  if (!f || std::memcmp(h.magic, "IEXACB1", 8) != 0 || h.src_len != key_text.size() || h.hash2 != fnv1a(key_text, 0x9e3779b97f4a7c15ull) ||
      h.cubin_len == 0 || h.cubin_len > (1ull << 30))
    return false;
  cubin.resize(h.cubin_len);
  f.read(cubin.data(), (std::streamsize)h.cubin_len);
  return (bool)f;
}
static void disk_cache_store(const std::string &key_text, const std::vector<char> &cubin) {
  const std::string dir = cache_dir();
  if (dir.empty() || cubin.empty()) return;
  char name[64], tmp[96];
  snprintf(name, sizeof name, "/%016llx.cubin", (unsigned long long)fnv1a(key_text, 1469598103934665603ull));
  snprintf(tmp, sizeof tmp, "%s.%ld.tmp", name, (long)getpid());
  CacheHeader h{};
  std::memcpy(h.magic, "IEXACB1", 8);
  h.src_len = key_text.size(); h.hash2 = fnv1a(key_text, 0x9e3779b97f4a7c15ull); h.cubin_len = cubin.size();
  {
    std::ofstream f(dir + tmp, std::ios::binary | std::ios::trunc);
    if (!f) return;
    f.write((const char *)&h, sizeof h);
    f.write(cubin.data(), (std::streamsize)cubin.size());
    if (!f) { f.close(); ::remove((dir + tmp).c_str()); return; }
  }
  if (::rename((dir + tmp).c_str(), (dir + name).c_str()) != 0) ::remove((dir + tmp).c_str());
}

static int g_nvrtc_compiles = 0, g_disk_hits = 0; // reporting (iexa_debug_cache_stats)
void cache_stats(int *compiles, int *disk_hits) { *compiles = g_nvrtc_compiles; *disk_hits = g_disk_hits; }

bool compile_cubin(const std::string &src, std::vector<char> &cubin, std::string &err) {
  DynApi &a = api();
  {
    std::lock_guard<std::mutex> g(cubin_cache_mutex());
    auto it = cubin_cache().find(src);
    if (it != cubin_cache().end() && !getenv("IEXA_DUMP_DIR")) { cubin = it->second; return true; }
  }
  const char *opts[] = {"--gpu-architecture=sm_100a", "-lineinfo", "--std=c++17", "-default-device"};
  std::string key_text = src;
  for (const char *o : opts) { key_text += '\n'; key_text += o; }
  { // the compiler that produced an image is part of its identity
    std::string verr;
    int maj = 0, min = 0;
    if (a.load_nvrtc(verr) && a.nvrtcVersion) a.nvrtcVersion(&maj, &min);
    key_text += "\nnvrtc " + std::to_string(maj) + "." + std::to_string(min);
  }
  if (!getenv("IEXA_DUMP_DIR") && disk_cache_load(key_text, cubin)) {
    std::lock_guard<std::mutex> g(cubin_cache_mutex());
    ++g_disk_hits;
    if (cubin_cache().size() >= 64) cubin_cache().clear();
    cubin_cache()[src] = cubin;
    return true;
  }
  if (!a.load_nvrtc(err)) return false;
  // optional dump so that `ncu --import-source on` can attach the generated translation unit
  std::string name = "iexa_generated.cu";
  if (const char *dir = getenv("IEXA_DUMP_DIR")) {
    size_t h = std::hash<std::string>()(src);
    char buf[64];
    snprintf(buf, sizeof buf, "/iexa_generated_%016zx.cu", h);
    name = std::string(dir) + buf;
    std::ofstream f(name);
    f << src;
  }
  nvrtcProgram prog = nullptr;
  if (a.nvrtcCreateProgram(&prog, src.c_str(), name.c_str(), 0, nullptr, nullptr) != 0) { err = "nvrtcCreateProgram failed"; return false; }
  int rc = a.nvrtcCompileProgram(prog, 4, opts);
  if (rc != 0) {
    size_t n = 0;
    a.nvrtcGetProgramLogSize(prog, &n);
    std::string log(n, '\0');
    if (n) a.nvrtcGetProgramLog(prog, &log[0]);
    err = "NVRTC compile failed (" + std::to_string(rc) + "): " + log.substr(0, 4000);
    a.nvrtcDestroyProgram(&prog);
    return false;
  }
  size_t n = 0;
  a.nvrtcGetCUBINSize(prog, &n);
  cubin.resize(n);
  a.nvrtcGetCUBIN(prog, cubin.data());
  a.nvrtcDestroyProgram(&prog);
  disk_cache_store(key_text, cubin);
  {
    std::lock_guard<std::mutex> g(cubin_cache_mutex());
    ++g_nvrtc_compiles;
    if (cubin_cache().size() >= 64) cubin_cache().clear();
    cubin_cache()[src] = cubin;
  }
  return true;
}

Specialiser::Specialiser() {}
Specialiser::~Specialiser() {
  for (void *m : module_) if (m && api().cuModuleUnload) api().cuModuleUnload((CUmodule)m);
  for (unsigned long long d : dev_allocs_) if (api().cuMemFree_v2) api().cuMemFree_v2(d);
}

static std::string cu_err(CUresult r) {
  const char *s = nullptr;
  api().cuGetErrorString(r, &s);
  return s ? s : "?";
}

bool Specialiser::build(Plan &plan, const std::vector<const void *> &col_dev_ptr, std::string &err) {
  return load_set(plan, col_dev_ptr, 0, true, err) && n_kernels_ > 0;
}
bool Specialiser::build_products(Plan &plan, const std::vector<const void *> &col_dev_ptr, std::string &err) {
  if (module_[1]) return true;
  return load_set(plan, col_dev_ptr, 1, false, err);
}
bool Specialiser::build_eval3(Plan &plan, const std::vector<const void *> &col_dev_ptr, std::string &err) {
  if (module_[2]) return true;
  return load_set(plan, col_dev_ptr, 2, false, err);
}

bool Specialiser::load_set(Plan &plan, const std::vector<const void *> &col_dev_ptr, int set, bool allow_regroup, std::string &err) {
  DynApi &a = api();
  if (!a.load_cuda(err)) return false;
  GeneratedSource gs = iexa::generate_source(plan, set);
  // Compile-time budget.  Models with thousands of small generators of a few shapes (large OPF
  // grids: one generator per bus / branch constraint) produce megabytes of straight-line code and
  // minutes of NVRTC time: same-shape generators are then canonicalised into one body with an
  // instance axis (shape classes); what still does not fit runs on the AOT interpreter kernels.
  size_t budget = 600 * 1024;
  if (const char *e = getenv("IEXA_CODEGEN_MAX_BYTES")) budget = (size_t)atoll(e);
  if (gs.text.size() > budget && !plan.class_mode_ && allow_regroup) {
    plan.build_groups(/*class_mode=*/true);
    gs = iexa::generate_source(plan, set);
  }
  if (gs.text.size() > budget) {
    err = "generated source (" + std::to_string(gs.text.size()) + " B) exceeds the specialisation budget (" +
          std::to_string(budget) + " B): using the interpreter kernels";
    return false;
  }
  std::vector<char> cubin;
  if (!compile_cubin(gs.text, cubin, err)) return false;
  CUmodule mod = nullptr;
  CUresult r = a.cuModuleLoadData(&mod, cubin.data());
  if (r != 0) { err = "cuModuleLoadData: " + cu_err(r); return false; }
  module_[set] = mod;
  // resolve the constant table
  std::vector<long long> vals(std::max<size_t>(gs.ci.size(), 1), 0);
  for (size_t i = 0; i < gs.ci.size(); ++i) {
    const CiEntry &e = gs.ci[i];
    if (e.kind == CI_SCHED) {
      const CbSchedule &S = gs.sched[e.a];
      vals[i] = e.b == 0 ? S.NB : e.b == 1 ? S.ntable
                : (e.b - 2 < (int)S.big.size() ? S.big[e.b - 2] : -1);
      continue;
    }
    const Group &G = plan.groups[e.group];
    const Iterator &itr = plan.itrs[G.itr];
    switch (e.kind) {
      case CI_SB: {
        const CbSchedule &S = gs.sched[e.a];
        vals[i] = spec_block();
        for (size_t j = 0; j < S.big.size(); ++j) if (S.big[j] == e.group) vals[i] = S.sb[j];
        break;
      }
      case CI_K0: vals[i] = G.k0; break;
      case CI_K1: vals[i] = G.k1; break;
      case CI_IDX_BASE: vals[i] = G.ctx.uidx[e.a].base; break;
      case CI_ICOL_PTR: vals[i] = (long long)(uintptr_t)col_dev_ptr[itr.int_cols[G.ctx.int_cols[e.a]].col]; break;
      case CI_FCOL_PTR: vals[i] = (long long)(uintptr_t)col_dev_ptr[itr.fp_cols[G.ctx.fp_cols[e.a]].col]; break;
      case CI_MEM_ROWLOC: vals[i] = plan.member(G, e.a).l0; break;
      case CI_CLS_NBLK: vals[i] = (G.k1 - G.k0 + spec_block() - 1) / spec_block(); break;
      case CI_CLS_ITAB: case CI_CLS_DTAB: {
        const std::vector<Generator> &gens = G.is_obj ? plan.objs : plan.cons;
        const size_t nis = G.ctx.uidx.size(), ncp = std::max<size_t>(G.cpar_nodes.size(), 1), ni = G.n_inst(), nm = G.members.size();
        const size_t W = nis + 4 * nm;
        std::vector<long long> it(ni * W);
        std::vector<double> dt(ni * ncp, 0.0);
        for (size_t q = 0; q < ni; ++q) {
          for (size_t u = 0; u < nis; ++u) it[q * W + u] = G.inst_base[q * nis + u];
          for (size_t j = 0; j < nm; ++j) {
            const Generator &g = gens[G.inst_gen(q, j)];
            it[q * W + nis + 4 * j] = g.l0;
            it[q * W + nis + 4 * j + 1] = g.l0;
            it[q * W + nis + 4 * j + 2] = g.l1;
            it[q * W + nis + 4 * j + 3] = g.l2;
          }
          for (size_t c = 0; c < G.cpar_nodes.size(); ++c) dt[q * ncp + c] = gens[G.inst_gen(q, G.cpar_member[c])].c.tape[G.cpar_nodes[c]].c;
        }
        const void *src = e.kind == CI_CLS_ITAB ? (const void *)it.data() : (const void *)dt.data();
        size_t bytes = e.kind == CI_CLS_ITAB ? it.size() * 8 : dt.size() * 8;
        CUdeviceptr d = 0;
        if ((r = a.cuMemAlloc_v2(&d, bytes ? bytes : 8)) != 0) { err = "cuMemAlloc(instance table): " + cu_err(r); return false; }
        dev_allocs_.push_back(d);
        if ((r = a.cuMemcpyHtoD_v2(d, src, bytes)) != 0) { err = "cuMemcpyHtoD(instance table): " + cu_err(r); return false; }
        vals[i] = (long long)d;
        break;
      }
      case CI_MEM_OUT: {
        const Generator &g = plan.member(G, e.a);
        vals[i] = (e.b == PROG_VAL || e.b == PROG_JV) ? g.l0 : e.b == PROG_D1 ? g.l1 : g.l2;
        break;
      }
    }
  }
  CUdeviceptr dptr = 0;
  size_t dbytes = 0;
  r = a.cuModuleGetGlobal_v2(&dptr, &dbytes, mod, "CI");
  if (r != 0) { err = "cuModuleGetGlobal(CI): " + cu_err(r); return false; }
  if (dbytes < vals.size() * 8) { err = "constant table size mismatch"; return false; }
  r = a.cuMemcpyHtoD_v2(dptr, vals.data(), vals.size() * 8);
  if (r != 0) { err = "cuMemcpyHtoD(CI): " + cu_err(r); return false; }
  for (int ks = 0; ks < KS__N; ++ks) {
    if (ks_set(ks) != set) continue;
    if (gs.groups_of[ks].empty()) continue;
    CUfunction f = nullptr;
    if (a.cuModuleGetFunction(&f, mod, kCbName[ks]) != 0 || !f) { err = std::string("missing kernel ") + kCbName[ks]; return false; }
    smem_[ks] = gs.smem_bytes[ks];
    if (smem_[ks] > 48 * 1024) {
      r = a.cuFuncSetAttribute(f, 8 /*CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES*/, (int)smem_[ks]);
      if (r != 0) { err = "cuFuncSetAttribute(max dynamic smem): " + cu_err(r); return false; }
    }
    fn_[ks] = f;
    groups_of_[ks] = gs.groups_of[ks];
    sched_[ks] = gs.sched[ks];
    ++n_kernels_;
  }
  return true;
}

bool Specialiser::launch(int ks, const WorkItem *work, const double *x, const double *theta,
                         const double *y, const double *v, double sigma, double *out, double *partials, cudaStream_t st,
                         std::string &err, double *out2, double *out3) {
  void *args[] = {(void *)&work, (void *)&x, (void *)&theta, (void *)&y, (void *)&v, (void *)&sigma, (void *)&out, (void *)&partials,
                  (void *)&out2, (void *)&out3};
  if (use_pdl() && api().cuLaunchKernelEx) {
    // CUlaunchConfig / CUlaunchAttribute of cuda.h (CUDA 12): the library does not link libcuda, so the two PODs are restated
    struct Attr { int id; char pad[4]; union { char pad[64]; int allowed; void *align_; } value; };
    struct Config { unsigned gx, gy, gz, bx, by, bz, smem; CUstream st; Attr *attrs; unsigned nattrs; };
    static_assert(sizeof(Attr) == 72, "CUlaunchAttribute layout");
    Attr at{};
    at.id = 6; // CU_LAUNCH_ATTRIBUTE_PROGRAMMATIC_STREAM_SERIALIZATION
    at.value.allowed = 1;
    Config cfg{(unsigned)sched_[ks].gx, (unsigned)sched_[ks].gy, 1, (unsigned)spec_block(), 1, 1, (unsigned)smem_[ks], (CUstream)st, &at, 1};
    CUresult r = api().cuLaunchKernelEx(&cfg, (CUfunction)fn_[ks], args, nullptr);
    if (r != 0) { err = "cuLaunchKernelEx: " + cu_err(r); return false; }
    return true;
  }
  CUresult r = api().cuLaunchKernel((CUfunction)fn_[ks], (unsigned)sched_[ks].gx, (unsigned)sched_[ks].gy, 1, (unsigned)spec_block(), 1, 1, (unsigned)smem_[ks], (CUstream)st, args, nullptr);
  if (r != 0) { err = "cuLaunchKernel: " + cu_err(r); return false; }
  return true;
}

} // namespace iexa

// engine.cu — sm_100a device engine: plan upload, the AOT tape-interpreter kernels, structure
// fill, reductions, and the launch logic of every NLPModels callback.
//
// Launch shape: ONE kernel launch per callback for the whole model (the reference stack
// launches one KernelAbstractions kernel per generator per callback plus fill!/compress
// kernels — SURVEY.md §2.2).  A work table maps blockIdx.x -> (generator, block of supports);
// every thread owns one support point k and runs the generator's register program.
// Outputs go straight to their precomputed slots  out[l + ostep*(k-k0) + c]  — every slot is
// rewritten on every call, so no fill!(vals, 0) pass is needed.
#include <cuda.h>          // types of the virtual-memory API only: the entry points are taken from libcuda.so.1 with dlsym
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <memory>

#include "codegen.hpp"
#include "engine.hpp"
#include "exec.hpp"

namespace iexa {

#define CK(call)                                                                  \
  do {                                                                            \
    cudaError_t e_ = (call);                                                      \
    if (e_ != cudaSuccess) {                                                      \
      err = std::string(#call) + ": " + cudaGetErrorString(e_);                   \
      return IEXA_ERR_CUDA;                                                       \
    }                                                                             \
  } while (0)

constexpr int BLOCK = 128;
enum { SINK_DENSE = 0, SINK_SUM = 1, SINK_SCATTER = 2, SINK_SCATTER_PROD = 3 };

// ------------------------------------------------------------------------------------------
// Tape interpreter.  Registers live in a per-thread local array (interleaved per thread by
// the hardware => coalesced, L1-resident for typical programs).  All control flow is
// warp-uniform: every thread of the block runs the same generator.
// ------------------------------------------------------------------------------------------
template <int NREG>
__global__ void __launch_bounds__(BLOCK)
interp_kernel(const GenD *__restrict__ gens, const WorkItem *__restrict__ work, int prog, int sink,
              const double *__restrict__ x, const double *__restrict__ theta,
              const double *__restrict__ y, const double *__restrict__ vec, double sigma, double *__restrict__ out,
              double *__restrict__ partials) {
  const WorkItem w = work[blockIdx.x];
  const GenD &g = gens[w.gen];
  long long k = g.k0 + (long long)w.blk * BLOCK + threadIdx.x;
  const bool active = k < g.k1;
  if (!active) k = g.k1 - 1; // keep loads in range; outputs are masked
  const ProgD P = g.prog[prog];
  const Instr *__restrict__ code = P.code;
  const double *__restrict__ cp = P.cpool;
  double r[NREG];
  double acc = 0.0;
  double W = sigma;
  if (!g.is_obj) W = y ? y[g.row_local + (k - g.k0)] : 0.0;
  const long long obase = g.out_local[prog] + (k - g.k0) * (long long)g.ostep[prog];

  for (int pc = 0; pc < P.ncode; ++pc) {
    const int4 raw = __ldg(reinterpret_cast<const int4 *>(code + pc));
    const int op = raw.x, dst = raw.y, ia = raw.z, ib = raw.w;
    switch (op) {
      case D_FIELD: r[dst] = col_fp(g.fcol[ia], k); break;
      case D_LOADX: r[dst] = __ldg(x + (idx_eval(g, ia, k) - 1)); break;
      case D_LOADP: r[dst] = __ldg(theta + (idx_eval(g, ia, k) - 1)); break;
      case D_LOADV: r[dst] = __ldg(vec + (idx_eval(g, ia, k) - 1)); break;
      case D_W: r[dst] = W; break;
      case D_SEL2: r[dst] = idx_eval(g, ia, k) == idx_eval(g, ib, k) ? 2.0 : 1.0; break;
      case D_SELNE: r[dst] = idx_eval(g, ia, k) != idx_eval(g, ib, k) ? 1.0 : 0.0; break;
      case D_OUT: {
        const double v = ia >= 0 ? r[ia] : cp[~ia];
        if (sink == SINK_DENSE) {
          if (active) out[obase + dst] = v;
        } else if (sink == SINK_SUM) {
          if (active) acc += v;
        } else if (sink == SINK_SCATTER_PROD) { // jtprod! / hprod!: out is zero-filled, every output adds
          const int islot = g.scat_slot[prog - PROG_JTV][dst];
          if (g.idx[islot].nterms == 0) {
            double s = active ? v : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
            if ((threadIdx.x & 31) == 0) atomicAdd(out + (g.idx[islot].base - 1), s);
          } else if (active) {
            atomicAdd(out + (idx_eval(g, islot, k) - 1), v);
          }
        } else {
          int islot = g.jac_slot[dst];
          if (islot < 0) { // single writer (Plan::analyse_grad): plain store, the range is not zero-filled
            if (active) out[idx_eval(g, ~islot, k) - 1] = v;
          } else if (g.idx[islot].nterms == 0) { // shared variable: warp-shuffle reduce, one atomic per warp
            double s = active ? v : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
            if ((threadIdx.x & 31) == 0) atomicAdd(out + (g.idx[islot].base - 1), s);
          } else if (active) {
            atomicAdd(out + (idx_eval(g, islot, k) - 1), v);
          }
        }
        break;
      }
      default: {
        const double a = ia >= 0 ? r[ia] : cp[~ia];
        const double b = ib >= 0 ? r[ib] : cp[~ib];
        r[dst] = eval_arith(op, a, b);
      }
    }
  }
  if (sink == SINK_SUM) {
    __shared__ double red[BLOCK / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < BLOCK / 32; ++i) s += red[i];
      partials[blockIdx.x] = s;
    }
  }
}

// same interpreter with the register file in global scratch (programs with > 256 live values:
// expanded measures, transform.jl:430-435).  scratch[reg * nthreads + tid].
__global__ void __launch_bounds__(BLOCK)
interp_kernel_big(const GenD *__restrict__ gens, const WorkItem *__restrict__ work, int prog, int sink,
                  const double *__restrict__ x, const double *__restrict__ theta,
                  const double *__restrict__ y, const double *__restrict__ vec, double sigma, double *__restrict__ out,
                  double *__restrict__ partials, double *__restrict__ scratch) {
  const WorkItem w = work[blockIdx.x];
  const GenD &g = gens[w.gen];
  long long k = g.k0 + (long long)w.blk * BLOCK + threadIdx.x;
  const bool active = k < g.k1;
  if (!active) k = g.k1 - 1;
  const ProgD P = g.prog[prog];
  const size_t nthr = (size_t)gridDim.x * BLOCK;
  double *r = scratch + ((size_t)blockIdx.x * BLOCK + threadIdx.x);
  double acc = 0.0;
  double W = sigma;
  if (!g.is_obj) W = y ? y[g.row_local + (k - g.k0)] : 0.0;
  const long long obase = g.out_local[prog] + (k - g.k0) * (long long)g.ostep[prog];
  for (int pc = 0; pc < P.ncode; ++pc) {
    const Instr I = P.code[pc];
    switch (I.op) {
      case D_FIELD: r[I.dst * nthr] = col_fp(g.fcol[I.a], k); break;
      case D_LOADX: r[I.dst * nthr] = x[idx_eval(g, I.a, k) - 1]; break;
      case D_LOADP: r[I.dst * nthr] = theta[idx_eval(g, I.a, k) - 1]; break;
      case D_LOADV: r[I.dst * nthr] = vec[idx_eval(g, I.a, k) - 1]; break;
      case D_W: r[I.dst * nthr] = W; break;
      case D_SEL2: r[I.dst * nthr] = idx_eval(g, I.a, k) == idx_eval(g, I.b, k) ? 2.0 : 1.0; break;
      case D_SELNE: r[I.dst * nthr] = idx_eval(g, I.a, k) != idx_eval(g, I.b, k) ? 1.0 : 0.0; break;
      case D_OUT: {
        const double v = I.a >= 0 ? r[I.a * nthr] : P.cpool[~I.a];
        if (sink == SINK_DENSE) { if (active) out[obase + I.dst] = v; }
        else if (sink == SINK_SUM) { if (active) acc += v; }
        else if (sink == SINK_SCATTER_PROD) { if (active) atomicAdd(out + (idx_eval(g, g.scat_slot[prog - PROG_JTV][I.dst], k) - 1), v); }
        else if (active) {
          const int is_ = g.jac_slot[I.dst];
          if (is_ < 0) out[idx_eval(g, ~is_, k) - 1] = v;
          else atomicAdd(out + (idx_eval(g, is_, k) - 1), v);
        }
        break;
      }
      default: {
        const double a = I.a >= 0 ? r[I.a * nthr] : P.cpool[~I.a];
        const double b = I.b >= 0 ? r[I.b * nthr] : P.cpool[~I.b];
        r[I.dst * nthr] = eval_arith(I.op, a, b);
      }
    }
  }
  if (sink == SINK_SUM) {
    __shared__ double red[BLOCK];
    red[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int i = 0; i < BLOCK; ++i) s += red[i];
      partials[blockIdx.x] = s;
    }
  }
}

// deterministic second stage of the objective reduction (fixed summation tree)
__global__ void __launch_bounds__(1024) reduce_partials_kernel(const double *__restrict__ p, int n,
                                                               double *__restrict__ out) {
  __shared__ double red[1024];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) s += p[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = red[0];
}

// Device-side transcription: fp iterator columns described by a closed form (plan.hpp: HostColumn::gen_*).  numpy's /
// InfiniteOpt's arithmetic restated operation by operation with explicitly rounded products and sums (no FMA
// contraction), so the column is bit-identical to the one the host path would have uploaded.
__device__ __forceinline__ double gen_pub(long long i, long long n, double a, double b, double step) {
  return i == n - 1 ? b : __dadd_rn(a, __dmul_rn((double)i, step));
}
// positions [j0, j1) of a K-long column (world > 1: the slice this rank reads); src / out are addressed by the GLOBAL position
__global__ void __launch_bounds__(256) gen_column_kernel(int kind, long long K, long long j0, long long j1, long long n, double a, double b, double step,
                                                         const double *__restrict__ src, double *__restrict__ out) {
  for (long long j = j0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; j < j1; j += (long long)gridDim.x * blockDim.x) {
    double v = 0.0;
    switch (kind) {
      case 1: v = n == 1 ? a : gen_pub(j, n, a, b, step); break;
      case 2: v = (j & 1) ? __dmul_rn(0.5, __dadd_rn(gen_pub(j >> 1, n, a, b, step), gen_pub((j >> 1) + 1, n, a, b, step))) : gen_pub(j >> 1, n, a, b, step); break;
      case 3: {
        double c = 0.0;
        if (j < K - 1) c = __ddiv_rn(__dsub_rn(src[j + 1], src[j]), 2.0);
        if (j > 0) c = __dadd_rn(c, __ddiv_rn(__dsub_rn(src[j], src[j - 1]), 2.0));
        v = c;
        break;
      }
      case 4: v = a; break;
    }
    out[j] = v;
  }
}

// zero-fill of the gradient entries no single-writer slot covers: ALL ranges in one launch (one cudaMemsetAsync per
// range is one stream operation each — four for the quadrotor, dozens for models with many variable blocks).
// ranges: [start, length] pairs (doubles); 16-byte stores where the alignment allows
__global__ void __launch_bounds__(256) zero_ranges_kernel(const long long *__restrict__ ranges, int nranges,
                                                          double *__restrict__ g) {
  for (int r = blockIdx.y; r < nranges; r += gridDim.y) {
    const long long s = ranges[2 * r], n = ranges[2 * r + 1];
    double *p = g + s;
    const long long head = (n > 0 && (((unsigned long long)p >> 3) & 1ull)) ? 1 : 0;
    const long long npair = (n - head) >> 1;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      if (head) p[0] = 0.0;
      if ((n - head) & 1) p[n - 1] = 0.0;
    }
    double2 *q = reinterpret_cast<double2 *>(p + head);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npair; i += (long long)gridDim.x * blockDim.x)
      q[i] = make_double2(0.0, 0.0);
  }
}

// rows/cols of the Jacobian (which=0) or lower-triangular Hessian (which=1); 1-based
template <typename IT>
__global__ void __launch_bounds__(BLOCK)
structure_kernel(const GenD *__restrict__ gens, const WorkItem *__restrict__ work, int which,
                 int local_rows, IT *__restrict__ rows, IT *__restrict__ cols) {
  const WorkItem w = work[blockIdx.x];
  const GenD &g = gens[w.gen];
  const long long k = g.k0 + (long long)w.blk * BLOCK + threadIdx.x;
  if (k >= g.k1) return;
  if (which >= 2) { // locality key of every COO slot: relative position of its support in the generator's range
    const int pr = which == 2 ? PROG_D1 : PROG_D2;
    const int n = g.ostep[pr];
    const long long base = g.out_local[pr] + (k - g.k0) * n;
    // buckets of 128 supports: long enough that a bucket's entries of one generator are a contiguous run of the COO
    // array, short enough that the runs of all generators of a bucket sit in cache together
    const long long nbk = (g.K + 127) >> 7;
    const IT key = (IT)(((k >> 7) << 20) / (nbk > 0 ? nbk : 1));
    for (int c = 0; c < n; ++c) rows[base + c] = key;
    return;
  }
  if (which == 0) {
    const int n = g.ostep[PROG_D1];
    const long long base = g.out_local[PROG_D1] + (k - g.k0) * n;
    const long long row = (local_rows ? g.row_local + (k - g.k0) : g.row_global + k) + 1;
    for (int c = 0; c < n; ++c) {
      rows[base + c] = (IT)row;
      cols[base + c] = (IT)idx_eval(g, g.jac_slot[c], k);
    }
  } else {
    const int n = g.ostep[PROG_D2];
    const long long base = g.out_local[PROG_D2] + (k - g.k0) * n;
    for (int c = 0; c < n; ++c) {
      const long long i = idx_eval(g, g.hess_slot[2 * c], k), j = idx_eval(g, g.hess_slot[2 * c + 1], k);
      rows[base + c] = (IT)(i >= j ? i : j);
      cols[base + c] = (IT)(i >= j ? j : i);
    }
  }
}

// ------------------------------------------------------------------------------------------
struct DevBuf {
  void *p = nullptr;
  size_t bytes = 0;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t ensure(size_t n) {
    if (n <= bytes) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; bytes = 0;
    cudaError_t e = cudaMalloc(&p, n ? n : 8);
    if (e == cudaSuccess) bytes = n;
    return e;
  }
  template <typename T> T *as() { return (T *)p; }
};

// A full-length-ADDRESSABLE device vector of which only some ranges are backed by memory (world > 1: theta is addressed by
// global parameter indices inside the kernels, but a rank reads only the slices of its own supports).  The CUDA virtual-memory
// API reserves the whole address range and maps physical granules (2 MB) over the ranges this rank reads; the kernels keep
// their global indices.  Small vectors, dense read sets, or a driver without the API: one ordinary allocation.
struct SparseDev {
  struct Vmm {
    CUresult (*GetGran)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*Reserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*Create)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
    CUresult (*Map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*SetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
    CUresult (*Unmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*Release)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*AddrFree)(CUdeviceptr, size_t) = nullptr;
    bool ok = false;
    Vmm() {
      void *h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
      if (!h) return;
      *(void **)&GetGran = dlsym(h, "cuMemGetAllocationGranularity");
      *(void **)&Reserve = dlsym(h, "cuMemAddressReserve");
      *(void **)&Create = dlsym(h, "cuMemCreate");
      *(void **)&Map = dlsym(h, "cuMemMap");
      *(void **)&SetAccess = dlsym(h, "cuMemSetAccess");
      *(void **)&Unmap = dlsym(h, "cuMemUnmap");
      *(void **)&Release = dlsym(h, "cuMemRelease");
      *(void **)&AddrFree = dlsym(h, "cuMemAddressFree");
      ok = GetGran && Reserve && Create && Map && SetAccess && Unmap && Release && AddrFree;
    }
  };
  static Vmm &vmm() { static Vmm v; return v; }

  DevBuf dense;
  CUdeviceptr va = 0;
  size_t va_size = 0, full = 0, resident = 0;
  std::vector<std::pair<size_t, size_t>> chunks;   // mapped byte ranges [lo, hi) of the reservation, sorted, disjoint
  std::vector<CUmemGenericAllocationHandle> handles;
  bool sparse = false;

  void *base() const { return sparse ? (void *)va : dense.p; }
  void release() {
    if (sparse) {
      for (size_t i = 0; i < chunks.size(); ++i) { vmm().Unmap(va + chunks[i].first, chunks[i].second - chunks[i].first); vmm().Release(handles[i]); }
      if (va) vmm().AddrFree(va, va_size);
    }
    chunks.clear(); handles.clear(); va = 0; va_size = 0; sparse = false; resident = 0;
  }
  ~SparseDev() { release(); }
  // ranges: byte ranges [lo, hi) that must be backed, sorted and disjoint
  cudaError_t alloc(size_t full_bytes, const std::vector<std::pair<size_t, size_t>> &ranges, int device, bool allow_sparse) {
    release();
    full = full_bytes;
    if (allow_sparse && vmm().ok && !getenv("IEXA_NO_VMM")) {
      CUmemAllocationProp prop{};
      prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
      prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
      prop.location.id = device;
      size_t G = 0;
      if (vmm().GetGran(&G, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM) == CUDA_SUCCESS && G > 0) {
        std::vector<std::pair<size_t, size_t>> ch;
        size_t tot = 0;
        for (auto &r : ranges) {
          if (r.second <= r.first) continue;
          size_t lo = r.first / G * G, hi = (r.second + G - 1) / G * G;
          if (!ch.empty() && lo <= ch.back().second) ch.back().second = std::max(ch.back().second, hi);
          else ch.emplace_back(lo, hi);
        }
        for (auto &c : ch) tot += c.second - c.first;
        const size_t vs = (full_bytes + G - 1) / G * G;
        // worth it only when the vector is large and the rank reads a minority of it
        if (full_bytes >= 8 * G && tot * 4 <= vs * 3 && !ch.empty()) {
          CUdeviceptr a = 0;
          if (vmm().Reserve(&a, vs, G, 0, 0) == CUDA_SUCCESS) {
            va = a; va_size = vs; sparse = true;
            CUmemAccessDesc acc{};
            acc.location = prop.location;
            acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
            bool good = true;
            for (auto &c : ch) {
              CUmemGenericAllocationHandle h;
              if (vmm().Create(&h, c.second - c.first, &prop, 0) != CUDA_SUCCESS) { good = false; break; }
              if (vmm().Map(va + c.first, c.second - c.first, 0, h, 0) != CUDA_SUCCESS) { vmm().Release(h); good = false; break; }
              chunks.push_back(c); handles.push_back(h);
              if (vmm().SetAccess(va + c.first, c.second - c.first, &acc, 1) != CUDA_SUCCESS) { good = false; break; }
            }
            if (good) { resident = tot; return cudaSuccess; }
            release();   // fall through to the ordinary allocation
          }
        }
      }
    }
    cudaError_t e = dense.ensure(full_bytes ? full_bytes : 8);
    if (e == cudaSuccess) resident = full_bytes;
    return e;
  }
  // host -> device copy of the bytes [off, off + n) of the FULL vector that are resident here (the rest belongs to other ranks)
  cudaError_t upload(const char *host_full, size_t off, size_t n, cudaStream_t st, bool async) {
    auto cp = [&](size_t lo, size_t hi) -> cudaError_t {
      if (hi <= lo) return cudaSuccess;
      char *d = (char *)base() + lo;
      return async ? cudaMemcpyAsync(d, host_full + (lo - off), hi - lo, cudaMemcpyHostToDevice, st)
                   : cudaMemcpy(d, host_full + (lo - off), hi - lo, cudaMemcpyHostToDevice);
    };
    if (!sparse) return cp(off, off + n);
    for (auto &c : chunks) {
      cudaError_t e = cp(std::max(off, c.first), std::min({off + n, c.second, full}));
      if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
  }
};

struct HostArena {
  std::vector<char> bytes;
  size_t add(const void *data, size_t n, size_t align = 16) {
    size_t off = (bytes.size() + align - 1) / align * align;
    bytes.resize(off + n);
    if (n) std::memcpy(bytes.data() + off, data, n);
    return off;
  }
};

class CudaEngine : public Engine {
 public:
  CudaEngine(Plan &plan, int device, uint32_t flags) : plan_(plan), device_(device), flags_(flags) {}
  ~CudaEngine() override {
    cudaSetDevice(device_);
    for (auto &kv : registered_) cudaHostUnregister(kv.first);
    if (pinned_f_) cudaFreeHost(pinned_f_);
    if (par_stage_) cudaFreeHost(par_stage_);
    if (par_stage_event_) cudaEventDestroy(par_stage_event_);
    if (stage_x_event_) cudaEventDestroy(stage_x_event_);
    spec_.reset();
  }

  int init(std::string &err) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
      cudaGetLastError();
      err = "no CUDA device available: the engine has no CPU fallback";
      return IEXA_ERR_CUDA;
    }
    if (device_ < 0 || device_ >= ndev) { err = "bad device ordinal"; return IEXA_ERR_INVALID; }
    CK(cudaSetDevice(device_));
    CK(cudaFree(0));
    int rc = upload(err);
    if (rc) return rc;
    CK(cudaMallocHost((void **)&pinned_f_, 64));
    if (!(flags_ & IEXA_F_NO_SPECIALISE)) {
      spec_.reset(new Specialiser());
      std::string serr;
      if (!spec_->build(plan_, col_dev_ptr_, serr) || build_group_tables(serr) != IEXA_OK) {
        // specialisation is an optimisation of the SAME device program; the AOT interpreter
        // kernels remain the (GPU) execution path.  Surface why.
        spec_error_ = serr;
        spec_.reset();
      }
    }
    return IEXA_OK;
  }

  // ---- callbacks ---------------------------------------------------------------------------
  int structure(int which, void *rows, void *cols, int idx_bytes, int memspace, void *stream,
                std::string &err) override {
    CK(cudaSetDevice(device_));
    cudaStream_t st = (cudaStream_t)stream;
    const bool locality = which >= 2; // 2 / 3: locality keys of the Jacobian / Hessian slots into `rows` (cols unused)
    const int64_t n = (which & 1) == 0 ? plan_.loc_nnzj : plan_.loc_nnzh;
    if (idx_bytes != 4 && idx_bytes != 8) { err = "idx_bytes must be 4 or 8"; return IEXA_ERR_INVALID; }
    void *dr = rows, *dc = cols;
    if (memspace == IEXA_MEM_HOST) {
      CK(stage_a_.ensure((size_t)n * idx_bytes));
      if (!locality) CK(stage_b_.ensure((size_t)n * idx_bytes));
      dr = stage_a_.p; dc = stage_b_.p;
    }
    const Table &T = table_[(which & 1) == 0 ? CB_JAC : CB_HESS];
    if (T.nblocks > 0) {
      if (idx_bytes == 4)
        structure_kernel<int32_t><<<T.nblocks, BLOCK, 0, st>>>(gens_dev_, T.work, which, 0, (int32_t *)dr, (int32_t *)dc);
      else
        structure_kernel<long long><<<T.nblocks, BLOCK, 0, st>>>(gens_dev_, T.work, which, 0, (long long *)dr, (long long *)dc);
      CK(cudaGetLastError());
    }
    if (memspace == IEXA_MEM_HOST) {
      CK(cudaMemcpyAsync(rows, dr, (size_t)n * idx_bytes, cudaMemcpyDeviceToHost, st));
      if (!locality) CK(cudaMemcpyAsync(cols, dc, (size_t)n * idx_bytes, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
    }
    return IEXA_OK;
  }

  int obj_device(const double *x_dev, double *f_dev, void *stream, std::string &err) override {
    CK(cudaSetDevice(device_));
    cudaStream_t st = (cudaStream_t)stream;
    const Table &T = table_[CB_OBJ];
    if (T.nblocks == 0) { CK(cudaMemsetAsync(f_dev, 0, 8, st)); return IEXA_OK; }
    const int nb = (spec_ && spec_->has(CB_OBJ)) ? gtable_[CB_OBJ].nblocks : T.nblocks;
    CK(partials_.ensure((size_t)std::max(nb, T.nblocks) * 8));
    int rc = launch(CB_OBJ, PROG_VAL, SINK_SUM, x_dev, nullptr, 1.0, nullptr, st, err);
    if (rc) return rc;
    reduce_partials_kernel<<<1, 1024, 0, st>>>(partials_.as<double>(), nb, f_dev);
    CK(cudaGetLastError());
    return IEXA_OK;
  }

  int obj(const double *x, double *f_host, int memspace, void *stream, std::string &err) override {
    CK(cudaSetDevice(device_));
    memspace = host_mode(memspace);
    cudaStream_t st = (cudaStream_t)stream;
    const double *xd = nullptr;
    int rc = in(x, plan_.nvar, memspace, stage_x_, st, xd, err);
    if (rc) return rc;
    CK(fdev_.ensure(8));
    rc = obj_device(xd, fdev_.as<double>(), stream, err);
    if (rc) return rc;
    CK(cudaMemcpyAsync(pinned_f_, fdev_.p, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *f_host = *pinned_f_;
    return IEXA_OK;
  }

  int grad(const double *x, double *g, int memspace, void *stream, std::string &err) override {
    CK(cudaSetDevice(device_));
    memspace = host_mode(memspace);
    cudaStream_t st = (cudaStream_t)stream;
    const double *xd = nullptr;
    int rc = in(x, plan_.nvar, memspace, stage_x_, st, xd, err);
    if (rc) return rc;
    double *gd = g;
    if (memspace == IEXA_MEM_HOST) { CK(stage_out_.ensure((size_t)plan_.nvar * 8)); gd = stage_out_.as<double>(); }
    // only the entries that are not written by a single-writer slot need zero-filling
    if ((rc = zero_ranges(plan_.grad_zero_ranges, zero_ranges_, gd, st, err))) return rc;
    rc = launch(CB_GRAD, PROG_D1, SINK_SCATTER, xd, nullptr, 1.0, gd, st, err);
    if (rc) return rc;
    return out(g, gd, plan_.nvar, memspace, st, err);
  }

  int cons(const double *x, double *c, int memspace, void *stream, std::string &err) override {
    return dense_cb(CB_CONS, PROG_VAL, x, nullptr, 1.0, c, plan_.loc_ncon, memspace, stream, err);
  }
  int jac(const double *x, double *vals, int memspace, void *stream, std::string &err) override {
    return dense_cb(CB_JAC, PROG_D1, x, nullptr, 1.0, vals, plan_.loc_nnzj, memspace, stream, err);
  }
  int hess(const double *x, const double *y, double sigma, double *vals, int memspace, void *stream,
           std::string &err) override {
    return dense_cb(CB_HESS, PROG_D2, x, y, sigma, vals, plan_.loc_nnzh, memspace, stream, err);
  }

  // cons! + jac_coord! + hess_coord! at one (x, y) in ONE launch (iexa_eval3): a constraint group evaluates value, first and
  // second order from one fused program — x / theta / columns are loaded once and the sin / cos of a state computed once
  // for all three — and the objective groups add their Hessian slots.  Three launches have three fill / drain phases:
  // on the reference's own study sizes (16 000 supports) and on 1/8 shards they are most of the time.  Compiled on first use;
  // falls back to the three separate callbacks when the fused kernel is not available.
  int eval3(const double *x, const double *y, double sigma, double *c, double *jvals, double *hvals, int memspace,
            void *stream, std::string &err) override {
    CK(cudaSetDevice(device_));
    if (spec_ && !eval3_tried_) {
      eval3_tried_ = true;
      std::string serr;
      if (!spec_->build_eval3(plan_, col_dev_ptr_, serr) || build_group_tables(serr, 2) != IEXA_OK) eval3_note_ = serr.empty() ? "unavailable" : serr;
    }
    const bool fused = spec_ && spec_->has(KS_EVAL3) && eval3_note_.empty() && gtable_[KS_EVAL3].nblocks > 0;
    if (!fused || memspace != IEXA_MEM_DEVICE) {
      int rc = cons(x, c, memspace, stream, err);
      if (rc) return rc;
      const int ms2 = memspace == IEXA_MEM_HOST ? (int)IEXA_MEM_HOST_SAME_X : memspace;
      if ((rc = jac(x, jvals, ms2, stream, err))) return rc;
      return hess(x, y, sigma, hvals, ms2, stream, err);
    }
    if (!spec_->launch(KS_EVAL3, gtable_[KS_EVAL3].work, x, theta_ptr(), y, nullptr, sigma, c, partials_.as<double>(),
                       (cudaStream_t)stream, err, jvals, hvals))
      return IEXA_ERR_CUDA;
    return IEXA_OK;
  }

  int jprod(const double *x, const double *v, double *Jv, int memspace, void *stream, std::string &err) override {
    return prod(CB_JPROD, x, nullptr, v, 1.0, Jv, memspace, stream, err);
  }
  int jtprod(const double *x, const double *v, double *Jtv, int memspace, void *stream, std::string &err) override {
    return prod(CB_JTPROD, x, nullptr, v, 1.0, Jtv, memspace, stream, err);
  }
  int hprod(const double *x, const double *y, const double *v, double sigma, double *Hv, int memspace,
            void *stream, std::string &err) override {
    return prod(CB_HPROD, x, y, v, sigma, Hv, memspace, stream, err);
  }

  // set_parameter! (infiniteopt_backend.jl:522,546): ordered on the CALLER's stream like every callback, so an update
  // can neither overtake a callback enqueued before it nor be overtaken by one enqueued after it.  The values are staged
  // in a pinned buffer of the engine (the caller's array may be pageable and may change right after the call).
  int set_par(int64_t off, int64_t n, const double *vals, void *stream, bool device_sync, std::string &err) override {
    CK(cudaSetDevice(device_));
    if (n <= 0) return IEXA_OK;
    if (device_sync) { // iexa_set_par: no stream given — safe against callbacks in flight on ANY stream
      CK(cudaDeviceSynchronize());
      CK(theta_.upload((const char *)vals, (size_t)off * 8, (size_t)n * 8, nullptr, false));
      return IEXA_OK;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (par_stage_event_) CK(cudaEventSynchronize(par_stage_event_)); // the previous update has left the staging buffer
    else CK(cudaEventCreateWithFlags(&par_stage_event_, cudaEventDisableTiming));
    if (par_stage_bytes_ < (size_t)n * 8) {
      if (par_stage_) cudaFreeHost(par_stage_);
      par_stage_ = nullptr; par_stage_bytes_ = 0;
      CK(cudaMallocHost((void **)&par_stage_, (size_t)n * 8));
      par_stage_bytes_ = (size_t)n * 8;
    }
    std::memcpy(par_stage_, vals, (size_t)n * 8);
    CK(theta_.upload((const char *)par_stage_, (size_t)off * 8, (size_t)n * 8, st, true));
    CK(cudaEventRecord(par_stage_event_, st));
    return IEXA_OK;
  }

  int launches(int cb) const override {
    if (cb < 0 || cb >= CB__N) return 0;
    if (cb >= CB_JPROD) {
      const Table &T = table_[cb == CB_JPROD ? CB_CONS : cb == CB_JTPROD ? CB_JAC : CB_HESS];
      if (cb == CB_JPROD) return T.nblocks > 0 ? 1 : 0;
      if (!(spec_ && spec_->products_built())) return 1 + (T.nblocks > 0 ? 1 : 0); // memset + interpreter
      const int w = cb == CB_JTPROD ? 0 : 1;
      return (plan_.scat_zero_ranges[w].empty() ? 0 : 1) + (spec_->has(w == 0 ? KS_JTPROD0 : KS_HPROD0) ? 1 : 0) +
             (spec_->has(w == 0 ? KS_JTPROD1 : KS_HPROD1) ? 1 : 0);
    }
    int n = table_[cb].nblocks > 0 ? 1 : 0;
    if (cb == CB_OBJ && n) n += 1;
    if (cb == CB_GRAD && n && !plan_.grad_zero_ranges.empty()) n += 1;
    return n;
  }
  int n_specialised() const override { return spec_ ? spec_->n_kernels() : 0; }
  const char *note() const override { return spec_error_.c_str(); }

 private:
  struct Table {
    WorkItem *work = nullptr;
    int nblocks = 0;
    int max_nreg = 0;
  };

  Plan &plan_;
  int device_;
  uint32_t flags_;
  DevBuf leaf_, desc_, gens_, work_;
  SparseDev theta_;                                      // full-length addressable; world > 1: only this rank's slices are backed
  std::vector<std::pair<int64_t, int64_t>> col_range_;   // resident positions [lo, hi) of every column (world > 1: the rank's slice)
  size_t col_bytes_ = 0, col_bytes_full_ = 0;
  double *theta_ptr() const { return (double *)theta_.base(); }
  DevBuf zero_ranges_;
  DevBuf partials_, fdev_, stage_x_, stage_y_, stage_v_, stage_out_, stage_a_, stage_b_, scratch_;
  std::vector<std::unique_ptr<DevBuf>> gen_bufs_; // device-generated iterator columns
  bool gen_failed_ = false;
  DevBuf scat_zero_[2];        // zero ranges of jtprod! / hprod! (Plan::scat_zero_ranges)
  int prod_max_nreg_[3] = {0, 0, 0}; // interpreter register file of the jv / jtv / hv programs
  bool prod_spec_tried_ = false, eval3_tried_ = false;
  std::string eval3_note_;
  double *par_stage_ = nullptr; size_t par_stage_bytes_ = 0; cudaEvent_t par_stage_event_ = nullptr;
  GenD *gens_dev_ = nullptr;
  std::vector<GenD> gens_host_; // device pointers inside; objs first then cons
  Table table_[CB__N];
  Table gtable_[KS__N]; // specialised path, per kernel slot: block -> (group, block of supports)
  DevBuf gwork_[3];
  std::vector<const void *> col_dev_ptr_;
  double *pinned_f_ = nullptr;
  std::map<void *, size_t> registered_;
  std::unique_ptr<Specialiser> spec_;
  std::string spec_error_;

  // ---- upload --------------------------------------------------------------------------------
  int upload(std::string &err) {
    Plan &P = plan_;
    HostArena A; // leaf data
    std::vector<size_t> col_off(P.columns.size(), (size_t)-1);
    col_dev_ptr_.assign(P.columns.size(), nullptr);
    std::vector<int32_t> gen_order; // generated columns in dependency order
    // world > 1: a rank keeps the positions [lo, hi) of a column that its own supports visit (Plan::column_read_ranges); the
    // device pointer is moved back by lo elements so that the kernels keep indexing with the global position
    col_range_ = P.column_read_ranges();
    if (P.world <= 1 || getenv("IEXA_NO_COLUMN_SLICES"))
      for (size_t c = 0; c < P.columns.size(); ++c) col_range_[c] = {0, P.columns[c].K};
    col_bytes_ = col_bytes_full_ = 0;
    std::vector<char> col_seen(P.columns.size(), 0);
    std::function<void(int32_t)> need_col = [&](int32_t c) {
      if (col_seen[c]) return;
      col_seen[c] = 1;
      const HostColumn &hc = P.columns[c];
      const int64_t lo = col_range_[c].first, n = std::max<int64_t>(col_range_[c].second - col_range_[c].first, 0);
      if (hc.is_int) {
        if (!hc.affine) { col_off[c] = A.add(hc.ivals.data() + lo, (size_t)n * 4); col_bytes_ += (size_t)n * 4; col_bytes_full_ += hc.ivals.size() * 4; }
      } else if (hc.gen_kind) { // generated on the device: own buffer, nothing uploaded
        if (hc.gen_src >= 0) need_col(hc.gen_src);
        gen_bufs_.emplace_back(new DevBuf());
        if (gen_bufs_.back()->ensure((size_t)(n > 0 ? n : 1) * 8) != cudaSuccess) { cudaGetLastError(); gen_failed_ = true; return; }
        col_dev_ptr_[c] = (char *)gen_bufs_.back()->p - (size_t)lo * 8;
        col_bytes_ += (size_t)n * 8; col_bytes_full_ += (size_t)hc.K * 8;
        gen_order.push_back(c);
      } else { col_off[c] = A.add(hc.fvals.data() + lo, (size_t)n * 8); col_bytes_ += (size_t)n * 8; col_bytes_full_ += hc.fvals.size() * 8; }
    };
    std::vector<Generator *> all;
    for (auto &g : P.objs) all.push_back(&g);
    for (auto &g : P.cons) all.push_back(&g);
    for (auto &g : P.pfuncs) all.push_back(&g); // parameter functions: value programs evaluated once, into theta
    struct Offs { size_t code[PROG__N], cpool[PROG__N], jac_slot, hess_slot, scat[2]; };
    std::vector<Offs> offs(all.size());
    for (size_t gi = 0; gi < all.size(); ++gi) {
      Generator &g = *all[gi];
      const Iterator &it = P.itrs[g.itr];
      for (int32_t s : g.c.int_cols) need_col(it.int_cols[s].col);
      for (int32_t s : g.c.fp_cols) need_col(it.fp_cols[s].col);
      Program *pr[PROG__N] = {&g.c.val, &g.c.d1, &g.c.d2, &g.c.jv, &g.c.jtv, &g.c.hv};
      for (int p = 0; p < PROG__N; ++p) {
        offs[gi].code[p] = A.add(pr[p]->code.data(), pr[p]->code.size() * sizeof(Instr));
        offs[gi].cpool[p] = A.add(pr[p]->cpool.data(), pr[p]->cpool.size() * 8);
      }
      offs[gi].scat[0] = A.add(g.c.jtv_slot.data(), g.c.jtv_slot.size() * 4);
      offs[gi].scat[1] = A.add(g.c.hv_slot.data(), g.c.hv_slot.size() * 4);
      if (!g.is_obj) { prod_max_nreg_[0] = std::max(prod_max_nreg_[0], g.c.jv.nreg); prod_max_nreg_[1] = std::max(prod_max_nreg_[1], g.c.jtv.nreg); }
      prod_max_nreg_[2] = std::max(prod_max_nreg_[2], g.c.hv.nreg);
      std::vector<int32_t> js(g.c.jac_slot);
      if (g.is_obj) for (size_t c = 0; c < js.size(); ++c) if (c < g.grad_direct.size() && g.grad_direct[c]) js[c] = ~js[c];
      offs[gi].jac_slot = A.add(js.data(), js.size() * 4);
      std::vector<int32_t> hs;
      for (auto &pr2 : g.c.hess_slot) { hs.push_back(pr2.first); hs.push_back(pr2.second); }
      offs[gi].hess_slot = A.add(hs.data(), hs.size() * 4);
    }
    if (gen_failed_) { err = "out of device memory for generated iterator columns"; return IEXA_ERR_CUDA; }
    CK(leaf_.ensure(A.bytes.size() + 16));
    CK(cudaMemcpy(leaf_.p, A.bytes.data(), A.bytes.size(), cudaMemcpyHostToDevice));
    char *lb = (char *)leaf_.p;
    for (size_t c = 0; c < P.columns.size(); ++c)
      if (col_off[c] != (size_t)-1) col_dev_ptr_[c] = lb + col_off[c] - (size_t)col_range_[c].first * (P.columns[c].is_int ? 4 : 8);
    for (int32_t c : gen_order) { // produce the generated columns on the device (a TRAPEZOID column after its source)
      const HostColumn &hc = P.columns[c];
      const double step = hc.gen_n > 1 ? (hc.gen_b - hc.gen_a) / (double)(hc.gen_n - 1) : 0.0;
      const int64_t j0 = col_range_[c].first, j1 = col_range_[c].second;
      if (j1 <= j0) continue;
      const int nb = (int)std::max<int64_t>(1, std::min<int64_t>((j1 - j0 + 255) / 256, 148 * 8));
      gen_column_kernel<<<nb, 256>>>(hc.gen_kind, hc.K, j0, j1, hc.gen_n, hc.gen_a, hc.gen_b, step,
                                     hc.gen_src >= 0 ? (const double *)col_dev_ptr_[hc.gen_src] : nullptr, (double *)col_dev_ptr_[c]);
      CK(cudaGetLastError());
    }

    HostArena D; // descriptors (ColD / IdxD arrays)
    struct DOffs { size_t icol, fcol, idx; };
    std::vector<DOffs> doffs(all.size());
    for (size_t gi = 0; gi < all.size(); ++gi) {
      Generator &g = *all[gi];
      const Iterator &it = P.itrs[g.itr];
      std::vector<ColD> ic, fc;
      for (int32_t s : g.c.int_cols) {
        const ColRef &r = it.int_cols[s];
        const HostColumn &hc = P.columns[r.col];
        ic.push_back(ColD{hc.affine ? nullptr : col_dev_ptr_[r.col], r.div, r.mod, hc.aa, hc.ab, hc.ac, hc.ad});
      }
      for (int32_t s : g.c.fp_cols) {
        const ColRef &r = it.fp_cols[s];
        fc.push_back(ColD{col_dev_ptr_[r.col], r.div, r.mod, 0, 0, 1, 0});
      }
      std::vector<IdxD> ix;
      for (auto &e : g.c.uidx) {
        IdxD d{};
        d.base = e.base;
        d.nterms = (int32_t)e.terms.size();
        for (int t = 0; t < d.nterms; ++t) { d.slot[t] = e.terms[t].first; d.coef[t] = e.terms[t].second; }
        ix.push_back(d);
      }
      doffs[gi].icol = D.add(ic.data(), ic.size() * sizeof(ColD));
      doffs[gi].fcol = D.add(fc.data(), fc.size() * sizeof(ColD));
      doffs[gi].idx = D.add(ix.data(), ix.size() * sizeof(IdxD));
    }
    CK(desc_.ensure(D.bytes.size() + 16));
    CK(cudaMemcpy(desc_.p, D.bytes.data(), D.bytes.size(), cudaMemcpyHostToDevice));
    char *db = (char *)desc_.p;

    gens_host_.assign(all.size(), GenD{});
    for (size_t gi = 0; gi < all.size(); ++gi) {
      Generator &g = *all[gi];
      GenD &d = gens_host_[gi];
      d.K = g.K; d.k0 = g.k0; d.k1 = g.k1;
      d.row_local = g.l0; d.row_global = g.o0;
      d.icol = (const ColD *)(db + doffs[gi].icol);
      d.fcol = (const ColD *)(db + doffs[gi].fcol);
      d.idx = (const IdxD *)(db + doffs[gi].idx);
      d.jac_slot = (const int32_t *)(lb + offs[gi].jac_slot);
      d.hess_slot = (const int32_t *)(lb + offs[gi].hess_slot);
      d.n_icol = (int32_t)g.c.int_cols.size();
      d.n_fcol = (int32_t)g.c.fp_cols.size();
      d.n_idx = (int32_t)g.c.uidx.size();
      d.is_obj = g.is_obj ? 1 : 0;
      Program *pr[PROG__N] = {&g.c.val, &g.c.d1, &g.c.d2, &g.c.jv, &g.c.jtv, &g.c.hv};
      d.scat_slot[0] = (const int32_t *)(lb + offs[gi].scat[0]);
      d.scat_slot[1] = (const int32_t *)(lb + offs[gi].scat[1]);
      for (int p = PROG_JV; p < PROG__N; ++p) { d.out_local[p] = g.l0; d.out_global[p] = g.o0; d.ostep[p] = 1; }
      for (int p = 0; p < PROG__N; ++p) {
        d.prog[p].code = (const Instr *)(lb + offs[gi].code[p]);
        d.prog[p].cpool = (const double *)(lb + offs[gi].cpool[p]);
        d.prog[p].ncode = (int32_t)pr[p]->code.size();
        d.prog[p].nreg = pr[p]->nreg;
        d.prog[p].nout = pr[p]->nout;
        d.prog[p].uses_w = pr[p]->uses_w;
      }
      d.out_local[PROG_VAL] = g.is_obj ? 0 : g.l0;
      d.out_local[PROG_D1] = g.is_obj ? 0 : g.l1;
      d.out_local[PROG_D2] = g.l2;
      d.out_global[PROG_VAL] = g.o0;
      d.out_global[PROG_D1] = g.is_obj ? g.og : g.o1;
      d.out_global[PROG_D2] = g.o2;
      d.ostep[PROG_VAL] = 1;
      d.ostep[PROG_D1] = g.c.o1step;
      d.ostep[PROG_D2] = g.c.o2step;
    }
    CK(gens_.ensure(gens_host_.size() * sizeof(GenD) + 16));
    CK(cudaMemcpy(gens_.p, gens_host_.data(), gens_host_.size() * sizeof(GenD), cudaMemcpyHostToDevice));
    gens_dev_ = gens_.as<GenD>();

    {
      std::vector<std::pair<size_t, size_t>> tr;
      for (auto &r : P.theta_read_ranges()) tr.emplace_back((size_t)r.first * 8, (size_t)r.second * 8);
      CK(theta_.alloc((size_t)(P.npar > 0 ? P.npar : 1) * 8, tr, device_, P.world > 1));
      if (P.npar > 0) CK(theta_.upload((const char *)P.theta.data(), 0, (size_t)P.npar * 8, nullptr, false));
    }

    // work tables
    const int nobj = (int)P.objs.size(), ncon = (int)P.cons.size();
    std::vector<WorkItem> items;
    size_t starts[6];
    auto add_range = [&](int g0, int g1, int prog, bool skip_empty_out, int &max_nreg) {
      for (int gi = g0; gi < g1; ++gi) {
        const Generator &g = *all[gi];
        const Program &pr = prog == PROG_VAL ? g.c.val : prog == PROG_D1 ? g.c.d1 : g.c.d2;
        if (skip_empty_out && pr.nout == 0) continue;
        max_nreg = std::max(max_nreg, pr.nreg);
        int64_t n = g.k1 - g.k0;
        for (int64_t b = 0; b * BLOCK < n; ++b) items.push_back(WorkItem{gi, (int32_t)b});
      }
    };
    int mr[5] = {0, 0, 0, 0, 0};
    starts[CB_OBJ] = items.size();  add_range(0, nobj, PROG_VAL, false, mr[CB_OBJ]);
    starts[CB_GRAD] = items.size(); add_range(0, nobj, PROG_D1, true, mr[CB_GRAD]);
    starts[CB_CONS] = items.size(); add_range(nobj, nobj + ncon, PROG_VAL, false, mr[CB_CONS]);
    starts[CB_JAC] = items.size();  add_range(nobj, nobj + ncon, PROG_D1, true, mr[CB_JAC]);
    starts[CB_HESS] = items.size(); add_range(0, nobj + ncon, PROG_D2, true, mr[CB_HESS]);
    starts[5] = items.size();
    CK(work_.ensure(items.size() * sizeof(WorkItem) + 16));
    if (!items.empty()) CK(cudaMemcpy(work_.p, items.data(), items.size() * sizeof(WorkItem), cudaMemcpyHostToDevice));
    for (int cb = 0; cb < 5; ++cb) {
      table_[cb].work = work_.as<WorkItem>() + starts[cb];
      table_[cb].nblocks = (int)(starts[cb + 1] - starts[cb]);
      table_[cb].max_nreg = mr[cb];
    }
    // parameter functions: theta blocks evaluated on the device, one after the other (a later one may read an earlier
    // one through PAR leaves), then mirrored into the host copy that iexa_get_par serves
    for (size_t pi = 0; pi < P.pfuncs.size(); ++pi) {
      const Generator &g = P.pfuncs[pi];
      const int gi = nobj + ncon + (int)pi;
      std::vector<WorkItem> pw;   // world > 1: the k-range whose theta entries this rank reads (Plan::shard_pfuncs)
      const int64_t nk = g.k1 - g.k0;
      for (int64_t b = 0; b * BLOCK < nk; ++b) pw.push_back(WorkItem{gi, (int32_t)b});
      if (pw.empty()) continue;
      DevBuf wb;
      CK(wb.ensure(pw.size() * sizeof(WorkItem)));
      CK(cudaMemcpy(wb.p, pw.data(), pw.size() * sizeof(WorkItem), cudaMemcpyHostToDevice));
      Table T;
      T.work = wb.as<WorkItem>(); T.nblocks = (int)pw.size(); T.max_nreg = g.c.val.nreg;
      int rc = launch_interp(T, g.c.val.nreg, PROG_VAL, SINK_DENSE, nullptr, nullptr, nullptr, 1.0, theta_ptr(), nullptr, err);
      if (rc) return rc;
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(P.theta.data() + g.l0, theta_ptr() + g.l0, (size_t)nk * 8, cudaMemcpyDeviceToHost));
    }
    return IEXA_OK;
  }

  // work tables of the specialised kernels of one module (set 0: the five callbacks; set 1: the products)
  int build_group_tables(std::string &err, int set = 0) {
    const int ks0 = set == 0 ? 0 : set == 1 ? KS_JPROD : KS_EVAL3, ks1 = set == 0 ? KS_JPROD : set == 1 ? KS_EVAL3 : KS__N;
    std::vector<WorkItem> items;
    size_t starts[KS__N + 1];
    for (int cb = ks0; cb < ks1; ++cb) {
      starts[cb] = items.size();
      if (!spec_->has(cb)) continue;
      // Big groups are scheduled arithmetically (CbSchedule, codegen.hpp); the table holds the rest.
      // Table blocks of different groups are interleaved by their RELATIVE position in the support
      // range: groups that walk the same supports then touch the same part of x at about the same
      // time, so the second reader hits in the 126 MB L2 instead of re-reading HBM.
      const CbSchedule &S = spec_->schedule(cb);
      struct Ord { double frac; int gi; int32_t b; };
      std::vector<Ord> ord;
      for (int gi : S.small) {
        const Group &G = plan_.groups[gi];
        int64_t n = G.k1 - G.k0;
        const int64_t SB = spec_block();
        int64_t nb = (n + SB - 1) / SB;
        const int64_t ninst = G.is_class ? ((int64_t)G.n_inst() + class_chunk() - 1) / class_chunk() : 1; // class groups: one block row per CHUNK of instances
        for (int64_t q = 0; q < ninst; ++q)
          for (int64_t b = 0; b < nb; ++b) ord.push_back(Ord{(b + 0.5) / (double)nb, gi, (int32_t)(q * nb + b)});
      }
      const char *oe = getenv("IEXA_ORDER");
      bool lpt = oe ? oe[0] == 'l' : false;
      if (lpt) { // heaviest groups first (longest processing time first): the kernel drains on short blocks
        const int prog = (cb == CB_OBJ || cb == CB_CONS) ? PROG_VAL : (cb == CB_GRAD || cb == CB_JAC) ? PROG_D1 : cb == CB_HESS ? PROG_D2
                         : cb == KS_JPROD ? PROG_JV : (cb == KS_JTPROD0 || cb == KS_JTPROD1) ? PROG_JTV : cb == KS_EVAL3 ? (int)GPROG_ALL : PROG_HV;
        std::stable_sort(ord.begin(), ord.end(), [&](const Ord &a, const Ord &b) {
          return plan_.groups[a.gi].prog[prog].code.size() > plan_.groups[b.gi].prog[prog].code.size();
        });
      } else if (oe ? oe[0] == 'g' : false) {
        // group-major: all blocks of one body back to back (fewer distinct case bodies resident at a time: the instruction
        // cache of a kernel with dozens of shape-class bodies — ncu: no_instruction is the 2nd stall of the 118-bus OPF cons)
        std::stable_sort(ord.begin(), ord.end(), [](const Ord &a, const Ord &b) { return a.gi < b.gi; });
      } else
      std::stable_sort(ord.begin(), ord.end(), [](const Ord &a, const Ord &b) { return a.frac < b.frac; });
      for (const Ord &o : ord) items.push_back(WorkItem{o.gi, o.b});
    }
    starts[ks1] = items.size();
    DevBuf &gw = gwork_[set];
    CK(gw.ensure(items.size() * sizeof(WorkItem) + 16));
    if (!items.empty()) CK(cudaMemcpy(gw.p, items.data(), items.size() * sizeof(WorkItem), cudaMemcpyHostToDevice));
    for (int cb = ks0; cb < ks1; ++cb) {
      gtable_[cb].work = gw.as<WorkItem>() + starts[cb];
      if (!spec_->has(cb)) { gtable_[cb].nblocks = 0; continue; }
      const CbSchedule &S = spec_->schedule(cb);
      if ((int64_t)(starts[cb + 1] - starts[cb]) != S.ntable) { err = "work table does not match the block schedule"; return IEXA_ERR_INVALID; }
      if (S.gy > 65535 || S.nblocks() > 0x7fffffffll) { err = "too many blocks for one launch"; return IEXA_ERR_INVALID; }
      gtable_[cb].nblocks = (int)S.nblocks(); // whole grid (arithmetic rows + table rows): one objective partial per block
    }
    return IEXA_OK;
  }

  // ---- helpers -------------------------------------------------------------------------------
  // Page-locking is EXPLICIT (iexa_host_register): the caller knows the lifetime of its vectors.  (Pinning
  // behind the caller's back is unsafe: freed-and-remapped host memory would keep a stale registration.)
 public:
  int jac_rowptr(void *rowptr, int idx_bytes, int memspace, void *stream, std::string &err) override {
    CK(cudaSetDevice(device_));
    if (!rowptr || (idx_bytes != 4 && idx_bytes != 8)) { err = "rowptr: null buffer or idx_bytes not 4/8"; return IEXA_ERR_INVALID; }
    if (!plan_.jac_is_csr()) { err = "the Jacobian COO order is not CSR (needs IEXA_SLOT_ORDER_JAC_ROW_SORTED and a static column order in every generator)"; return IEXA_ERR_STATE; }
    if (idx_bytes == 4 && plan_.loc_nnzj > INT32_MAX) { err = "nnzj does not fit Int32 row pointers"; return IEXA_ERR_INVALID; }
    const std::vector<int64_t> rp = plan_.jac_rowptr();
    const size_t n = rp.size();
    std::vector<int32_t> rp32;
    const void *src = rp.data();
    if (idx_bytes == 4) { rp32.assign(rp.begin(), rp.end()); src = rp32.data(); }
    if (memspace == IEXA_MEM_HOST) { std::memcpy(rowptr, src, n * (size_t)idx_bytes); return IEXA_OK; }
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemcpyAsync(rowptr, src, n * (size_t)idx_bytes, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));   // the staging vectors die with this call
    return IEXA_OK;
  }
  void device_bytes(int64_t *out) const override {
    size_t gen = 0;
    for (auto &b : gen_bufs_) gen += b->bytes;
    out[0] = (int64_t)col_bytes_; out[1] = (int64_t)col_bytes_full_;
    out[2] = (int64_t)(plan_.npar > 0 ? theta_.resident : 0); out[3] = (int64_t)plan_.npar * 8;
    out[4] = (int64_t)(leaf_.bytes + gen + desc_.bytes + gens_.bytes + work_.bytes + zero_ranges_.bytes + partials_.bytes + fdev_.bytes +
                       scat_zero_[0].bytes + scat_zero_[1].bytes + gwork_[0].bytes + gwork_[1].bytes + gwork_[2].bytes + scratch_.bytes) - out[0];
    out[5] = (int64_t)(stage_x_.bytes + stage_y_.bytes + stage_v_.bytes + stage_out_.bytes + stage_a_.bytes + stage_b_.bytes);
  }
  int get_column(int32_t col, double *out_host, std::string &err) override {
    CK(cudaSetDevice(device_));
    if (col < 0 || col >= (int32_t)col_dev_ptr_.size() || !col_dev_ptr_[col] || plan_.columns[col].is_int) { err = "column is not resident on the device"; return IEXA_ERR_INVALID; }
    const int64_t lo = col_range_[col].first, hi = col_range_[col].second;   // world > 1: only this rank's slice is resident
    if (hi > lo) CK(cudaMemcpy(out_host + lo, (const double *)col_dev_ptr_[col] + lo, (size_t)(hi - lo) * 8, cudaMemcpyDeviceToHost));
    return IEXA_OK;
  }
  int host_register(void *p, size_t bytes, std::string &err) override {
    CK(cudaSetDevice(device_));
    if (!p || bytes == 0) { err = "null buffer"; return IEXA_ERR_INVALID; }
    if (registered_.count(p)) return IEXA_OK;
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type != cudaMemoryTypeUnregistered) return IEXA_OK; // already pinned
    cudaGetLastError();
    CK(cudaHostRegister(p, bytes, cudaHostRegisterDefault));
    registered_[p] = bytes;
    return IEXA_OK;
  }
  int host_unregister(void *p, std::string &err) override {
    CK(cudaSetDevice(device_));
    auto it = registered_.find(p);
    if (it == registered_.end()) return IEXA_OK;
    CK(cudaHostUnregister(p));
    registered_.erase(it);
    return IEXA_OK;
  }

 private:
  int in(const double *src, int64_t n, int memspace, DevBuf &stage, cudaStream_t st, const double *&dev,
         std::string &err) {
    if (!src) { dev = nullptr; return IEXA_OK; }
    if (memspace == IEXA_MEM_DEVICE) { dev = src; return IEXA_OK; }
    const bool is_x = &stage == &stage_x_;
    if (is_x && same_x_ && stage_x_valid_ && stage.bytes >= (size_t)n * 8) { // new_x == false
      // the device copy may have been uploaded on ANOTHER stream: order this call after that upload
      if (st != stage_x_stream_ && stage_x_event_) CK(cudaStreamWaitEvent(st, stage_x_event_, 0));
      dev = stage.as<double>();
      return IEXA_OK;
    }
    CK(stage.ensure((size_t)n * 8));
    if (is_x && plan_.world > 1) {
      // sharded: only the parts of x this rank READS cross PCIe (own supports of every block, shared variables, halos);
      // the rest of the staging buffer is never addressed by this rank's kernels
      if (x_ranges_.empty()) x_ranges_ = plan_.x_read_ranges();
      for (auto &r : x_ranges_)
        CK(cudaMemcpyAsync(stage.as<double>() + (r.first - 1), src + (r.first - 1), (size_t)(r.second - r.first + 1) * 8, cudaMemcpyHostToDevice, st));
    } else
    CK(cudaMemcpyAsync(stage.p, src, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    if (is_x) {
      stage_x_valid_ = true;
      if (!stage_x_event_) CK(cudaEventCreateWithFlags(&stage_x_event_, cudaEventDisableTiming));
      CK(cudaEventRecord(stage_x_event_, st));
      stage_x_stream_ = st;
    }
    dev = stage.as<double>();
    return IEXA_OK;
  }
  // IEXA_MEM_HOST_SAME_X is IEXA_MEM_HOST plus the promise that x has not changed since the last host call
  int host_mode(int memspace) {
    same_x_ = memspace == IEXA_MEM_HOST_SAME_X;
    return same_x_ ? (int)IEXA_MEM_HOST : memspace;
  }
  bool same_x_ = false, stage_x_valid_ = false;
  std::vector<std::pair<int64_t, int64_t>> x_ranges_; // 1-based inclusive (Plan::x_read_ranges)
  cudaEvent_t stage_x_event_ = nullptr; cudaStream_t stage_x_stream_ = nullptr;
  int out(double *dst, const double *dev, int64_t n, int memspace, cudaStream_t st, std::string &err) {
    if (memspace == IEXA_MEM_DEVICE) return IEXA_OK;
    CK(cudaMemcpyAsync(dst, dev, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return IEXA_OK;
  }

  int launch(int cb, int prog, int sink, const double *xd, const double *yd, double sigma, double *outd,
             cudaStream_t st, std::string &err) {
    const Table &T = table_[cb];
    if (T.nblocks == 0) return IEXA_OK;
    double *part = partials_.as<double>();
    const double *th = theta_ptr();
    if (spec_ && spec_->has(cb)) {
      const Table &GT = gtable_[cb];
      if (GT.nblocks == 0) return IEXA_OK;
      if (!spec_->launch(cb, GT.work, xd, th, yd, nullptr, sigma, outd, part, st, err)) return IEXA_ERR_CUDA;
      return IEXA_OK;
    }
    return launch_interp(T, T.max_nreg, prog, sink, xd, yd, nullptr, sigma, outd, st, err);
  }

  int launch_interp(const Table &T, int max_nreg, int prog, int sink, const double *xd, const double *yd, const double *vd,
                    double sigma, double *outd, cudaStream_t st, std::string &err) {
    if (T.nblocks == 0) return IEXA_OK;
    double *part = partials_.as<double>();
    const double *th = theta_ptr();
    if (max_nreg <= 32)
      interp_kernel<32><<<T.nblocks, BLOCK, 0, st>>>(gens_dev_, T.work, prog, sink, xd, th, yd, vd, sigma, outd, part);
    else if (max_nreg <= 96)
      interp_kernel<96><<<T.nblocks, BLOCK, 0, st>>>(gens_dev_, T.work, prog, sink, xd, th, yd, vd, sigma, outd, part);
    else if (max_nreg <= 256)
      interp_kernel<256><<<T.nblocks, BLOCK, 0, st>>>(gens_dev_, T.work, prog, sink, xd, th, yd, vd, sigma, outd, part);
    else {
      CK(scratch_.ensure((size_t)max_nreg * T.nblocks * BLOCK * 8));
      interp_kernel_big<<<T.nblocks, BLOCK, 0, st>>>(gens_dev_, T.work, prog, sink, xd, th, yd, vd, sigma, outd, part,
                                                     scratch_.as<double>());
    }
    CK(cudaGetLastError());
    return IEXA_OK;
  }

  int dense_cb(int cb, int prog, const double *x, const double *y, double sigma, double *o, int64_t n,
               int memspace, void *stream, std::string &err) {
    CK(cudaSetDevice(device_));
    memspace = host_mode(memspace);
    cudaStream_t st = (cudaStream_t)stream;
    const double *xd = nullptr, *yd = nullptr;
    int rc = in(x, plan_.nvar, memspace, stage_x_, st, xd, err);
    if (rc) return rc;
    rc = in(y, plan_.loc_ncon, memspace, stage_y_, st, yd, err);
    if (rc) return rc;
    double *od = o;
    if (memspace == IEXA_MEM_HOST) { CK(stage_out_.ensure((size_t)n * 8)); od = stage_out_.as<double>(); }
    rc = launch(cb, prog, SINK_DENSE, xd, yd, sigma, od, st, err);
    if (rc) return rc;
    return out(o, od, n, memspace, st, err);
  }

  int zero_ranges(const std::vector<std::pair<int64_t, int64_t>> &ranges, DevBuf &tab, double *dst, cudaStream_t st, std::string &err) {
    if (ranges.empty()) return IEXA_OK;
    const int nr = (int)ranges.size();
    if (!tab.p) {
      std::vector<long long> rr;
      for (auto &zr : ranges) { rr.push_back(zr.first); rr.push_back(zr.second); }
      CK(tab.ensure(rr.size() * 8));
      CK(cudaMemcpy(tab.p, rr.data(), rr.size() * 8, cudaMemcpyHostToDevice));
    }
    int64_t longest = 0;
    for (auto &zr : ranges) longest = std::max<int64_t>(longest, zr.second);
    const int gx = (int)std::max<int64_t>(1, std::min<int64_t>((longest / 2 + 255) / 256, 148 * 8));
    zero_ranges_kernel<<<dim3(gx, std::min(nr, 65535)), 256, 0, st>>>(tab.as<long long>(), nr, dst);
    CK(cudaGetLastError());
    return IEXA_OK;
  }

  // jprod! / jtprod! / hprod! — fused product kernels: the first / second order programs with a product epilogue
  // (gen.hpp: build_products), no COO values are materialised.
  //   jprod!:  Jv[row] = sum_c d1_c * v[col_c] in registers, one coalesced store per row; no atomics (bit-reproducible);
  //   jtprod! / hprod!: per group ONE sum per touched variable (contributions of all fused rows combined in registers);
  //   phase-0 groups store their single-writer blocks directly (no zero-fill), phase 1 adds the rest with atomics
  //   (warp-shuffle reduced for shared variables) — Plan::analyse_scatter.
  int prod(int cb, const double *x, const double *y, const double *v, double sigma, double *o, int memspace,
           void *stream, std::string &err) {
    CK(cudaSetDevice(device_));
    memspace = host_mode(memspace);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nv = cb == CB_JTPROD ? plan_.loc_ncon : plan_.nvar;
    const int64_t no = cb == CB_JPROD ? plan_.loc_ncon : plan_.nvar;
    if (spec_ && !prod_spec_tried_) { // the product kernels are compiled on their first use (second NVRTC module)
      prod_spec_tried_ = true;
      std::string serr;
      if (!spec_->build_products(plan_, col_dev_ptr_, serr) || build_group_tables(serr, 1) != IEXA_OK) { prod_note_ = serr; spec_error_ = "product kernels: " + serr; }
    }
    const double *xd = nullptr, *yd = nullptr, *vd = nullptr;
    int rc;
    if ((rc = in(x, plan_.nvar, memspace, stage_x_, st, xd, err))) return rc;
    if ((rc = in(y, plan_.loc_ncon, memspace, stage_y_, st, yd, err))) return rc;
    if ((rc = in(v, nv, memspace, stage_v_, st, vd, err))) return rc;
    double *od = o;
    if (memspace == IEXA_MEM_HOST) { CK(stage_out_.ensure((size_t)no * 8)); od = stage_out_.as<double>(); }
    const bool spec = spec_ && spec_->products_built() && prod_note_.empty();
    double *part = partials_.as<double>();
    const double *th = theta_ptr();
    if (cb == CB_JPROD) {
      if (spec) {
        if (spec_->has(KS_JPROD) && gtable_[KS_JPROD].nblocks > 0 &&
            !spec_->launch(KS_JPROD, gtable_[KS_JPROD].work, xd, th, nullptr, vd, 1.0, od, part, st, err)) return IEXA_ERR_CUDA;
      } else if ((rc = launch_interp(table_[CB_CONS], prod_max_nreg_[0], PROG_JV, SINK_DENSE, xd, nullptr, vd, 1.0, od, st, err))) return rc;
      return out(o, od, no, memspace, st, err);
    }
    const int w = cb == CB_JTPROD ? 0 : 1;
    // J'v: the root weight leaf of every row is v[row] — the kernels read it through their y argument
    const double *wy = w == 0 ? vd : yd, *wv = w == 0 ? nullptr : vd;
    if (spec) {
      if ((rc = zero_ranges(plan_.scat_zero_ranges[w], scat_zero_[w], od, st, err))) return rc;
      for (int ph = 0; ph < 2; ++ph) {
        const int ks = (w == 0 ? KS_JTPROD0 : KS_HPROD0) + ph;
        if (spec_->has(ks) && gtable_[ks].nblocks > 0 &&
            !spec_->launch(ks, gtable_[ks].work, xd, th, wy, wv, sigma, od, part, st, err)) return IEXA_ERR_CUDA;
      }
    } else {
      CK(cudaMemsetAsync(od, 0, (size_t)no * 8, st));
      if ((rc = launch_interp(table_[w == 0 ? CB_JAC : CB_HESS], prod_max_nreg_[1 + w], w == 0 ? PROG_JTV : PROG_HV, SINK_SCATTER_PROD,
                              xd, wy, wv, sigma, od, st, err))) return rc;
    }
    return out(o, od, no, memspace, st, err);
  }
  std::string prod_note_;
};

Engine *make_cuda_engine(Plan &plan, int device, uint32_t flags, std::string &err) {
  std::unique_ptr<CudaEngine> e(new CudaEngine(plan, device, flags));
  if (e->init(err) != IEXA_OK) return nullptr;
  return e.release();
}

} // namespace iexa

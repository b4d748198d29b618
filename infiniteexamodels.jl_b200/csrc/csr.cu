// csr.cu — COO -> CSR value permutation with duplicate summation.
//
// MadNLPGPU assembles the KKT matrix from the COO Jacobian/Hessian values with a `transfer!`
// kernel before every cuDSS refactorisation (caller side of
// ext/InfiniteExaModelsMadNLP.jl:49-50).  Setup (once): sort the 1-based COO pattern by
// (row, col), collapse duplicates into CSR entries and remember, for every CSR entry, the
// segment of the sorted permutation that feeds it.  Apply (every iteration): one thread per
// CSR entry gathers and sums its segment — no atomics, bit-reproducible, reads each COO value
// exactly once and writes each CSR value exactly once.
//
// Work order.  Walking the CSR entries in CSR order reads the COO values with the stride of the
// evaluator's per-support tiles (o2step doubles): the sources of one CSR row block are 1 double out
// of every 32-byte sector, and the other doubles of that sector belong to row blocks that are
// processed much later — every sector came from HBM ~4 times (Hessian of the 10^6-support
// quadrotor: 0.80 ms, 39 % of roofline).  The threads therefore walk the CSR entries ordered by the
// COO position of their FIRST source: all entries fed by the same support tile are handled
// together, the COO values are read as a stream (each sector once), and only the 8-byte result
// is scattered — to piecewise-contiguous runs that the L2 merges before write-back.  When the
// pattern has no duplicates the source of thread t is simply coo[t].
#include <cuda_runtime.h>
#include <thrust/binary_search.h>
#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/scan.h>
#include <thrust/sequence.h>
#include <thrust/sort.h>

#include <string>

#include "../../include/iexa.h"

struct iexa_csr {
  int device = 0;
  int64_t nrows = 0, ncols = 0, nnz = 0, csr_nnz = 0;
  int32_t *perm = nullptr;     // [nnz]   sources of the CSR entries, grouped per entry IN WORK ORDER (COO positions)
  int32_t *seg = nullptr;      // [csr_nnz+1] segment starts in perm, in work order
  int32_t *order = nullptr;    // [csr_nnz] work position -> CSR entry
  bool nodup = false;          // csr_nnz == nnz: the source of work position t is coo[t]
  int32_t *rowptr = nullptr;   // [nrows+1]
  int32_t *colind = nullptr;   // [csr_nnz]
  double *stage_in = nullptr, *stage_out = nullptr; // host-memspace staging
};

namespace iexa { extern thread_local std::string g_last_error; }
#define g_csr_err iexa::g_last_error
namespace {

template <typename IT>
__global__ void make_keys(int64_t n, const IT *rows, const IT *cols, int64_t ncols, int64_t *keys) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    keys[i] = ((int64_t)rows[i] - 1) * ncols + ((int64_t)cols[i] - 1);
}
__global__ void head_flags(int64_t n, const int64_t *keys, int64_t *flag) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    flag[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}
__global__ void fill_heads(int64_t n, const int64_t *keys, const int64_t *pos, int64_t ncols, int32_t *seg,
                           int32_t *colind, int64_t *ukeys) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (i == 0 || keys[i] != keys[i - 1]) {
      int64_t p = pos[i] - 1; // inclusive scan
      seg[p] = (int32_t)i;
      colind[p] = (int32_t)(keys[i] % ncols);
      ukeys[p] = keys[i];
    }
  }
}
__global__ void row_starts(int64_t nrows, int64_t ncols, const int64_t *ukeys, int64_t nu, int32_t *rowptr) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= nrows; r += (int64_t)gridDim.x * blockDim.x) {
    int64_t target = r * ncols, lo = 0, hi = nu; // first unique key >= target
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (ukeys[mid] < target) lo = mid + 1; else hi = mid; }
    rowptr[r] = (int32_t)lo;
  }
}
// setup helpers for the work order
// work-order key of a CSR entry: (locality key of its first source, COO position of its first source)
__global__ void first_source(int64_t nu, const int32_t *seg, const int32_t *perm, const int32_t *loc, int64_t *first) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nu; p += (int64_t)gridDim.x * blockDim.x) {
    const int32_t src = perm[seg[p]];
    first[p] = ((int64_t)(loc ? loc[src] : 0) << 32) | (int64_t)(uint32_t)src;
  }
}
__global__ void seg_lengths(int64_t nu, const int32_t *seg, const int32_t *order, int32_t *len) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nu; t += (int64_t)gridDim.x * blockDim.x) {
    const int32_t p = order[t];
    len[t] = seg[p + 1] - seg[p];
  }
}
__global__ void regroup_sources(int64_t nu, const int32_t *seg, const int32_t *order, const int32_t *perm,
                                const int32_t *seg2, int32_t *perm2) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nu; t += (int64_t)gridDim.x * blockDim.x) {
    const int32_t p = order[t], b = seg[p], n = seg[p + 1] - b, o = seg2[t];
    for (int32_t i = 0; i < n; ++i) perm2[o + i] = perm[b + i];
  }
}
// U independent (segment -> sources -> values) chains per thread: one chain is three dependent global loads, and a
// single chain per thread keeps too few bytes in flight (measured: 3.1 TB/s of metadata + value traffic)
template <int U>
__global__ void __launch_bounds__(256)
csr_apply_kernel(int64_t csr_nnz, const int32_t *__restrict__ seg, const int32_t *__restrict__ perm,
                 const int32_t *__restrict__ order, const double *__restrict__ coo, double *__restrict__ csr) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t0 < csr_nnz; t0 += U * stride) {
    int32_t b[U], e[U], dst[U];
    double s[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t t = t0 + u * stride;
      const bool ok = t < csr_nnz;
      b[u] = ok ? __ldg(seg + t) : 0;
      e[u] = ok ? __ldg(seg + t + 1) : 0;
      dst[u] = ok ? __ldg(order + t) : -1;
    }
    int32_t p[U];
#pragma unroll
    for (int u = 0; u < U; ++u) p[u] = b[u] < e[u] ? __ldg(perm + b[u]) : -1;
#pragma unroll
    for (int u = 0; u < U; ++u) s[u] = p[u] >= 0 ? __ldg(coo + p[u]) : 0.0;
#pragma unroll
    for (int u = 0; u < U; ++u)
      for (int32_t q = b[u] + 1; q < e[u]; ++q) s[u] += __ldg(coo + __ldg(perm + q)); // sources in COO order: fixed summation order
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (dst[u] >= 0) csr[dst[u]] = s[u];
  }
}
template <int U>
__global__ void __launch_bounds__(256)
csr_apply_nodup_kernel(int64_t nnz, const int32_t *__restrict__ order, const double *__restrict__ coo, double *__restrict__ csr) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t0 < nnz; t0 += U * stride) {
    int32_t dst[U];
    double v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t t = t0 + u * stride;
      dst[u] = t < nnz ? __ldg(order + t) : -1;
      v[u] = t < nnz ? __ldg(coo + t) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (dst[u] >= 0) csr[dst[u]] = v[u];
  }
}
int grid_for(int64_t n) { int64_t b = (n + 255) / 256; return (int)(b < 1 ? 1 : (b > 148 * 32 ? 148 * 32 : b)); }
} // namespace

#define CCK(call)                                                                     \
  do {                                                                                \
    cudaError_t e_ = (call);                                                          \
    if (e_ != cudaSuccess) { g_csr_err = std::string(#call) + ": " + cudaGetErrorString(e_); return IEXA_ERR_CUDA; } \
  } while (0)

// temporaries of the setup: freed on every exit path (an out-of-memory error at config-3 sizes must not leave several
// GB of device memory behind on the device the solver needs)
struct TmpBuf {
  void *p = nullptr;
  ~TmpBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 8); }
  void release() { if (p) cudaFree(p); p = nullptr; }
  template <typename T> T *as() { return (T *)p; }
};
struct CsrOwner { // owns the handle (and through iexa_csr_destroy everything it holds) until the setup has succeeded
  iexa_csr *h;
  ~CsrOwner() { if (h) iexa_csr_destroy(h); }
};

extern "C" {

int32_t iexa_csr_create(iexa_csr **out, int64_t nrows, int64_t ncols, int64_t nnz, const void *rows,
                        const void *cols, int32_t idx_bytes, int32_t memspace, int32_t device) {
  return iexa_csr_create_keyed(out, nrows, ncols, nnz, rows, cols, idx_bytes, nullptr, memspace, device);
}

int32_t iexa_csr_create_keyed(iexa_csr **out, int64_t nrows, int64_t ncols, int64_t nnz, const void *rows,
                              const void *cols, int32_t idx_bytes, const int32_t *keys_in, int32_t memspace,
                              int32_t device) {
  if (!out || nnz < 0 || nrows < 0 || ncols <= 0 || (idx_bytes != 4 && idx_bytes != 8)) { g_csr_err = "bad arguments"; return IEXA_ERR_INVALID; }
  *out = nullptr;
  if (nnz >= (1ll << 31)) { g_csr_err = "nnz exceeds int32 permutation range"; return IEXA_ERR_UNSUPPORTED; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); g_csr_err = "no CUDA device (no CPU fallback)"; return IEXA_ERR_CUDA; }
  CCK(cudaSetDevice(device));
  CsrOwner own{new iexa_csr()};
  iexa_csr *h = own.h;
  h->device = device; h->nrows = nrows; h->ncols = ncols; h->nnz = nnz;
  TmpBuf dr, dc, keys, flag, ukeys, seg0b, first, dloc, perm2;
  const void *r = rows, *c = cols;
  if (memspace == IEXA_MEM_HOST && nnz > 0) {
    CCK(dr.alloc((size_t)nnz * idx_bytes)); CCK(dc.alloc((size_t)nnz * idx_bytes));
    CCK(cudaMemcpy(dr.p, rows, (size_t)nnz * idx_bytes, cudaMemcpyHostToDevice));
    CCK(cudaMemcpy(dc.p, cols, (size_t)nnz * idx_bytes, cudaMemcpyHostToDevice));
    r = dr.p; c = dc.p;
  }
  size_t n1 = (size_t)(nnz > 0 ? nnz : 1);
  CCK(keys.alloc(n1 * 8)); CCK(flag.alloc(n1 * 8));
  CCK(cudaMalloc(&h->perm, n1 * 4));
  try {
    if (nnz > 0) {
      if (idx_bytes == 4) make_keys<int32_t><<<grid_for(nnz), 256>>>(nnz, (const int32_t *)r, (const int32_t *)c, ncols, keys.as<int64_t>());
      else make_keys<long long><<<grid_for(nnz), 256>>>(nnz, (const long long *)r, (const long long *)c, ncols, keys.as<int64_t>());
      thrust::sequence(thrust::device, h->perm, h->perm + nnz);
      thrust::stable_sort_by_key(thrust::device, keys.as<int64_t>(), keys.as<int64_t>() + nnz, h->perm);
      head_flags<<<grid_for(nnz), 256>>>(nnz, keys.as<int64_t>(), flag.as<int64_t>());
      thrust::inclusive_scan(thrust::device, flag.as<int64_t>(), flag.as<int64_t>() + nnz, flag.as<int64_t>());
      CCK(cudaMemcpy(&h->csr_nnz, flag.as<int64_t>() + (nnz - 1), 8, cudaMemcpyDeviceToHost));
    }
    size_t nu = (size_t)(h->csr_nnz > 0 ? h->csr_nnz : 1);
    CCK(seg0b.alloc((nu + 1) * 4)); // segment starts in (row, col)-sorted order
    int32_t *seg0 = seg0b.as<int32_t>();
    CCK(cudaMalloc(&h->colind, nu * 4));
    CCK(ukeys.alloc(nu * 8));
    CCK(cudaMalloc(&h->rowptr, (size_t)(nrows + 1) * 4));
    if (nnz > 0) {
      fill_heads<<<grid_for(nnz), 256>>>(nnz, keys.as<int64_t>(), flag.as<int64_t>(), ncols, seg0, h->colind, ukeys.as<int64_t>());
      const int32_t nnz32 = (int32_t)nnz;
      CCK(cudaMemcpy(seg0 + h->csr_nnz, &nnz32, 4, cudaMemcpyHostToDevice));
    } else {
      int32_t z = 0;
      CCK(cudaMemcpy(seg0, &z, 4, cudaMemcpyHostToDevice));
    }
    row_starts<<<grid_for(nrows + 1), 256>>>(nrows, ncols, ukeys.as<int64_t>(), h->csr_nnz, h->rowptr);
    CCK(cudaDeviceSynchronize());
    keys.release(); flag.release(); ukeys.release();
    // work order: CSR entries sorted by the COO position of their first source
    const int64_t nuu = h->csr_nnz;
    CCK(cudaMalloc(&h->order, nu * 4));
    h->nodup = nuu == nnz && !keys_in; // with locality keys the work order is not the COO order
    if (nuu > 0) {
      const int32_t *loc = keys_in;
      if (keys_in && memspace == IEXA_MEM_HOST) {
        CCK(dloc.alloc(n1 * 4));
        CCK(cudaMemcpy(dloc.p, keys_in, (size_t)nnz * 4, cudaMemcpyHostToDevice));
        loc = dloc.as<int32_t>();
      }
      CCK(first.alloc(nu * 8));
      first_source<<<grid_for(nuu), 256>>>(nuu, seg0, h->perm, loc, first.as<int64_t>());
      thrust::sequence(thrust::device, h->order, h->order + nuu);
      thrust::stable_sort_by_key(thrust::device, first.as<int64_t>(), first.as<int64_t>() + nuu, h->order);
      first.release(); dloc.release();
      if (!h->nodup) {
        CCK(cudaMalloc(&h->seg, (nu + 1) * 4));
        CCK(perm2.alloc(n1 * 4));
        seg_lengths<<<grid_for(nuu), 256>>>(nuu, seg0, h->order, h->seg);
        CCK(cudaMemset(h->seg + nuu, 0, 4));
        thrust::exclusive_scan(thrust::device, h->seg, h->seg + nuu + 1, h->seg);
        regroup_sources<<<grid_for(nuu), 256>>>(nuu, seg0, h->order, h->perm, h->seg, perm2.as<int32_t>());
        CCK(cudaDeviceSynchronize());
        cudaFree(h->perm);
        h->perm = perm2.as<int32_t>();
        perm2.p = nullptr; // ownership moved into the handle
      } else {
        CCK(cudaDeviceSynchronize());
        cudaFree(h->perm); h->perm = nullptr; // identity in work order
      }
    }
  } catch (const std::exception &e) { // thrust reports allocation failures of its temporaries by throwing
    cudaGetLastError();
    g_csr_err = std::string("COO->CSR setup: ") + e.what();
    return IEXA_ERR_CUDA;
  }
  own.h = nullptr;
  *out = h;
  return IEXA_OK;
}

int32_t iexa_csr_destroy(iexa_csr *h) {
  if (!h) return IEXA_OK;
  cudaSetDevice(h->device);
  cudaFree(h->perm); cudaFree(h->seg); cudaFree(h->order); cudaFree(h->rowptr); cudaFree(h->colind);
  if (h->stage_in) cudaFree(h->stage_in);
  if (h->stage_out) cudaFree(h->stage_out);
  delete h;
  return IEXA_OK;
}

int64_t iexa_csr_nnz(const iexa_csr *h) { return h ? h->csr_nnz : -1; }

int32_t iexa_csr_pattern(const iexa_csr *h, int32_t *rowptr, int32_t *colind, int32_t memspace) {
  if (!h) { g_csr_err = "null handle"; return IEXA_ERR_INVALID; }
  CCK(cudaSetDevice(h->device));
  cudaMemcpyKind kind = memspace == IEXA_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  CCK(cudaMemcpy(rowptr, h->rowptr, (size_t)(h->nrows + 1) * 4, kind));
  if (h->csr_nnz > 0) CCK(cudaMemcpy(colind, h->colind, (size_t)h->csr_nnz * 4, kind));
  return IEXA_OK;
}

int32_t iexa_csr_apply(iexa_csr *h, const double *coo_vals, double *csr_vals, int32_t memspace, void *stream) {
  if (!h) { g_csr_err = "null handle"; return IEXA_ERR_INVALID; }
  CCK(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const double *in = coo_vals;
  double *outp = csr_vals;
  if (memspace == IEXA_MEM_HOST) {
    if (!h->stage_in) CCK(cudaMalloc(&h->stage_in, (size_t)(h->nnz > 0 ? h->nnz : 1) * 8));
    if (!h->stage_out) CCK(cudaMalloc(&h->stage_out, (size_t)(h->csr_nnz > 0 ? h->csr_nnz : 1) * 8));
    CCK(cudaMemcpyAsync(h->stage_in, coo_vals, (size_t)h->nnz * 8, cudaMemcpyHostToDevice, st));
    in = h->stage_in; outp = h->stage_out;
  }
  if (h->csr_nnz > 0) {
    if (h->nodup) csr_apply_nodup_kernel<4><<<grid_for((h->csr_nnz + 3) / 4), 256, 0, st>>>(h->csr_nnz, h->order, in, outp);
    else csr_apply_kernel<1><<<grid_for(h->csr_nnz), 256, 0, st>>>(h->csr_nnz, h->seg, h->perm, h->order, in, outp);
    CCK(cudaGetLastError());
  }
  if (memspace == IEXA_MEM_HOST) {
    CCK(cudaMemcpyAsync(csr_vals, outp, (size_t)h->csr_nnz * 8, cudaMemcpyDeviceToHost, st));
    CCK(cudaStreamSynchronize(st));
  }
  return IEXA_OK;
}

} // extern "C"

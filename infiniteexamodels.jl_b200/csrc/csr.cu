// csr.cu — COO -> CSR value permutation with duplicate summation.
//
// MadNLPGPU assembles the KKT matrix from the COO Jacobian/Hessian values with a `transfer!`
// kernel before every cuDSS refactorisation (caller side of
// ext/InfiniteExaModelsMadNLP.jl:49-50).  Setup (once): sort the 1-based COO pattern by
// (row, col), collapse duplicates into CSR entries and remember, for every CSR entry, the
// segment of the sorted permutation that feeds it.  Apply (every iteration): one thread per
// CSR entry gathers and sums its segment — no atomics, bit-reproducible, reads each COO value
// exactly once and writes each CSR value exactly once.
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
  int32_t *perm = nullptr;     // [nnz]   sorted position -> COO position
  int64_t *seg = nullptr;      // [csr_nnz+1] segment starts in sorted order
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
__global__ void fill_heads(int64_t n, const int64_t *keys, const int64_t *pos, int64_t ncols, int64_t *seg,
                           int32_t *colind, int64_t *ukeys) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (i == 0 || keys[i] != keys[i - 1]) {
      int64_t p = pos[i] - 1; // inclusive scan
      seg[p] = i;
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
__global__ void __launch_bounds__(256)
csr_apply_kernel(int64_t csr_nnz, const int64_t *__restrict__ seg, const int32_t *__restrict__ perm,
                 const double *__restrict__ coo, double *__restrict__ csr) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < csr_nnz; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = seg[p], e = seg[p + 1];
    double s = 0.0;
    for (int64_t q = b; q < e; ++q) s += __ldg(coo + perm[q]);
    csr[p] = s;
  }
}
int grid_for(int64_t n) { int64_t b = (n + 255) / 256; return (int)(b < 1 ? 1 : (b > 148 * 32 ? 148 * 32 : b)); }
} // namespace

#define CCK(call)                                                                     \
  do {                                                                                \
    cudaError_t e_ = (call);                                                          \
    if (e_ != cudaSuccess) { g_csr_err = std::string(#call) + ": " + cudaGetErrorString(e_); return IEXA_ERR_CUDA; } \
  } while (0)

extern "C" {

int32_t iexa_csr_create(iexa_csr **out, int64_t nrows, int64_t ncols, int64_t nnz, const void *rows,
                        const void *cols, int32_t idx_bytes, int32_t memspace, int32_t device) {
  if (!out || nnz < 0 || nrows < 0 || ncols <= 0 || (idx_bytes != 4 && idx_bytes != 8)) { g_csr_err = "bad arguments"; return IEXA_ERR_INVALID; }
  if (nnz >= (1ll << 31)) { g_csr_err = "nnz exceeds int32 permutation range"; return IEXA_ERR_UNSUPPORTED; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); g_csr_err = "no CUDA device (no CPU fallback)"; return IEXA_ERR_CUDA; }
  CCK(cudaSetDevice(device));
  iexa_csr *h = new iexa_csr();
  h->device = device; h->nrows = nrows; h->ncols = ncols; h->nnz = nnz;
  void *dr = nullptr, *dc = nullptr;
  const void *r = rows, *c = cols;
  if (memspace == IEXA_MEM_HOST && nnz > 0) {
    CCK(cudaMalloc(&dr, (size_t)nnz * idx_bytes)); CCK(cudaMalloc(&dc, (size_t)nnz * idx_bytes));
    CCK(cudaMemcpy(dr, rows, (size_t)nnz * idx_bytes, cudaMemcpyHostToDevice));
    CCK(cudaMemcpy(dc, cols, (size_t)nnz * idx_bytes, cudaMemcpyHostToDevice));
    r = dr; c = dc;
  }
  int64_t *keys = nullptr, *flag = nullptr, *ukeys = nullptr;
  size_t n1 = (size_t)(nnz > 0 ? nnz : 1);
  CCK(cudaMalloc(&keys, n1 * 8)); CCK(cudaMalloc(&flag, n1 * 8));
  CCK(cudaMalloc(&h->perm, n1 * 4));
  if (nnz > 0) {
    if (idx_bytes == 4) make_keys<int32_t><<<grid_for(nnz), 256>>>(nnz, (const int32_t *)r, (const int32_t *)c, ncols, keys);
    else make_keys<long long><<<grid_for(nnz), 256>>>(nnz, (const long long *)r, (const long long *)c, ncols, keys);
    thrust::sequence(thrust::device, h->perm, h->perm + nnz);
    thrust::stable_sort_by_key(thrust::device, keys, keys + nnz, h->perm);
    head_flags<<<grid_for(nnz), 256>>>(nnz, keys, flag);
    thrust::inclusive_scan(thrust::device, flag, flag + nnz, flag);
    CCK(cudaMemcpy(&h->csr_nnz, flag + (nnz - 1), 8, cudaMemcpyDeviceToHost));
  }
  size_t nu = (size_t)(h->csr_nnz > 0 ? h->csr_nnz : 1);
  CCK(cudaMalloc(&h->seg, (nu + 1) * 8));
  CCK(cudaMalloc(&h->colind, nu * 4));
  CCK(cudaMalloc(&ukeys, nu * 8));
  CCK(cudaMalloc(&h->rowptr, (size_t)(nrows + 1) * 4));
  if (nnz > 0) {
    fill_heads<<<grid_for(nnz), 256>>>(nnz, keys, flag, ncols, h->seg, h->colind, ukeys);
    CCK(cudaMemcpy(h->seg + h->csr_nnz, &nnz, 8, cudaMemcpyHostToDevice));
  } else {
    int64_t z = 0;
    CCK(cudaMemcpy(h->seg, &z, 8, cudaMemcpyHostToDevice));
  }
  row_starts<<<grid_for(nrows + 1), 256>>>(nrows, ncols, ukeys, h->csr_nnz, h->rowptr);
  CCK(cudaDeviceSynchronize());
  cudaFree(keys); cudaFree(flag); cudaFree(ukeys);
  if (dr) cudaFree(dr);
  if (dc) cudaFree(dc);
  *out = h;
  return IEXA_OK;
}

int32_t iexa_csr_destroy(iexa_csr *h) {
  if (!h) return IEXA_OK;
  cudaSetDevice(h->device);
  cudaFree(h->perm); cudaFree(h->seg); cudaFree(h->rowptr); cudaFree(h->colind);
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
    csr_apply_kernel<<<grid_for(h->csr_nnz), 256, 0, st>>>(h->csr_nnz, h->seg, h->perm, in, outp);
    CCK(cudaGetLastError());
  }
  if (memspace == IEXA_MEM_HOST) {
    CCK(cudaMemcpyAsync(csr_vals, outp, (size_t)h->csr_nnz * 8, cudaMemcpyDeviceToHost, st));
    CCK(cudaStreamSynchronize(st));
  }
  return IEXA_OK;
}

} // extern "C"

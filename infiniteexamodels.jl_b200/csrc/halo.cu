// halo.cu — x halo exchange and small all-reduce over NVLink peer memory (one process per GPU, CUDA IPC).
//
// Sharded evaluation: every rank owns a part of the iterate x and READS a little more — the
// replicated finite / shared variables and the shard-boundary halos of shifted references (y[i-1] of finite
// differences, lower-bound / internal nodes of collocation elements): a few dozen doubles per neighbour.  Sending those
// through NCCL point-to-point costs index-gather kernels, a group launch and index-scatter kernels — 0.12 ms on 8 GPUs
// for 180 doubles, more than the sharded evaluation itself (0.07 ms).  Here every rank maps its peers' x buffers and a
// small flag block into its address space once (cudaIpc*), and one exchange is ONE small kernel per rank:
//
//   1. ready : tell every owner I receive from that my x is ready to take the halo of THIS epoch (stream order: all my
//              callbacks of the previous iterate, and whatever rewrote my part of x since, precede this kernel);
//   2. push  : for every reader of mine — wait for its ready flag, store my boundary values straight into ITS x at their global
//              positions (peer stores over NVLink), __threadfence_system(), raise my flag in its flag block;
//   3. wait  : until every owner I receive from has raised its flag for this epoch.
//
// The callbacks that follow on the stream read x as usual.  No host synchronisation, no staging copies, no index
// kernels.  The small all-reduce (objective value + the shared gradient slice: a handful of doubles) works the same way:
// every rank stores its partial vector into slot [rank] of every peer's landing zone (double-buffered by epoch parity),
// raises a flag, waits for all flags and sums the world's partials in rank order — deterministic, and bit-identical on
// all ranks.  Larger slices go through NCCL (dist.py).
//
// Spin loops are bounded (≈ 4 s): a missing peer sets an error word instead of hanging the GPU.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstring>
#include <string>
#include <vector>

#include "../../include/iexa.h"

namespace iexa { extern thread_local std::string g_last_error; }

namespace {

constexpr int RED_MAX = 1024;      // doubles per rank of the small all-reduce
constexpr int MAX_WORLD = 64;

// flag block of one rank (device memory, mapped by every peer)
struct Block {
  unsigned long long data_flag[MAX_WORLD];      // [src]    src has pushed its halo of epoch (value) into my x
  unsigned long long ack_flag[MAX_WORLD];       // [reader] reader's x is ready to receive the halo of epoch (value)
  unsigned long long red_flag[2][MAX_WORLD];    // [parity][src]
  unsigned long long status;                    // != 0: a bounded wait expired
  double red_buf[2][MAX_WORLD][RED_MAX];
};

struct SendRange { int peer; long long lo, len; };

struct Dev {                      // kernel argument (by value)
  Block *blk[MAX_WORLD];          // every rank's flag block in MY address space (own = local)
  double *x[MAX_WORLD];           // every rank's x in MY address space (own = unused)
  const SendRange *sends; int nsend;
  const int *send_peers; int nsend_peers;   // distinct readers of mine
  const int *recv_peers; int nrecv_peers;   // distinct owners I receive from
  int rank, world;
};

__device__ __forceinline__ unsigned long long ld_flag(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_flag(unsigned long long *p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// bounded spin: false when the flag did not arrive within ~4 s
__device__ bool wait_ge(const unsigned long long *p, unsigned long long want, unsigned long long *status, unsigned long long code) {
  const long long t0 = clock64();
  while (ld_flag(p) < want) {
    __nanosleep(64);
    if (clock64() - t0 > 8000000000ll) { st_flag(status, code); return false; }
  }
  return true;
}

__global__ void __launch_bounds__(256) halo_exchange_kernel(Dev d, const double *__restrict__ x, unsigned long long epoch) {
  Block *me = d.blk[d.rank];
  const int tid = threadIdx.x;
  // 1. ready flags (nothing is waited for before them: no circular wait between ranks)
  if (tid < d.nrecv_peers) st_flag(&d.blk[d.recv_peers[tid]]->ack_flag[d.rank], epoch);
  // 2. pushes: one reader after the other (a reader gets a few dozen doubles)
  __shared__ int ok;
  for (int sp = 0; sp < d.nsend_peers; ++sp) {
    const int peer = d.send_peers[sp];
    if (tid == 0) ok = wait_ge(&me->ack_flag[peer], epoch, &me->status, 1000 + peer) ? 1 : 0;
    __syncthreads();
    if (ok) {
      double *__restrict__ px = d.x[peer];
      for (int s = 0; s < d.nsend; ++s) {
        if (d.sends[s].peer != peer) continue;
        const long long lo = d.sends[s].lo, n = d.sends[s].len;
        for (long long i = tid; i < n; i += blockDim.x) px[lo + i] = x[lo + i];
      }
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0) st_flag(&d.blk[peer]->data_flag[d.rank], epoch);
  }
  // 3. my halos
  if (tid < d.nrecv_peers) wait_ge(&me->data_flag[d.recv_peers[tid]], epoch, &me->status, 2000 + d.recv_peers[tid]);
  __syncthreads();
}

// buf[0..n) <- sum over ranks, summed in rank order on every rank
__global__ void __launch_bounds__(256) small_allreduce_kernel(Dev d, double *__restrict__ buf, int n, unsigned long long epoch) {
  const int tid = threadIdx.x, par = (int)(epoch & 1ull);
  Block *me = d.blk[d.rank];
  for (int p = 0; p < d.world; ++p) {
    double *dst = d.blk[p]->red_buf[par][d.rank];
    for (int i = tid; i < n; i += blockDim.x) dst[i] = buf[i];
  }
  __threadfence_system();
  __syncthreads();
  if (tid < d.world) st_flag(&d.blk[tid]->red_flag[par][d.rank], epoch);
  if (tid < d.world) wait_ge(&me->red_flag[par][tid], epoch, &me->status, 3000 + tid);
  __syncthreads();
  for (int i = tid; i < n; i += blockDim.x) {
    double s = 0.0;
    for (int p = 0; p < d.world; ++p) s += __ldcv(&me->red_buf[par][p][i]);
    buf[i] = s;
  }
}

typedef int (*cuMemGetAddressRange_t)(unsigned long long *, size_t *, unsigned long long);

} // namespace

struct iexa_halo {
  int device = 0, rank = 0, world = 1;
  Block *local = nullptr;
  Dev dev{};
  std::vector<void *> opened;       // cudaIpcOpenMemHandle results (to close)
  std::vector<SendRange> sends;
  std::vector<int> send_peers, recv_peers;
  void *d_sends = nullptr, *d_send_peers = nullptr, *d_recv_peers = nullptr;
  bool tables_dirty = true;
  unsigned long long epoch = 0, red_epoch = 0;
  const void *x_exported = nullptr;
};

#define HFAIL(code, msg) do { iexa::g_last_error = (msg); return (code); } while (0)
#define HCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { iexa::g_last_error = std::string(#call) + ": " + cudaGetErrorString(e_); return IEXA_ERR_CUDA; } } while (0)

extern "C" {

int32_t iexa_halo_create(iexa_halo **out, int32_t device, int32_t rank, int32_t world) {
  if (!out || world < 1 || world > MAX_WORLD || rank < 0 || rank >= world) HFAIL(IEXA_ERR_INVALID, "halo: bad rank / world");
  HCK(cudaSetDevice(device));
  iexa_halo *h = new iexa_halo();
  h->device = device; h->rank = rank; h->world = world;
  cudaError_t e = cudaMalloc((void **)&h->local, sizeof(Block));
  if (e != cudaSuccess) { delete h; HFAIL(IEXA_ERR_CUDA, std::string("halo: cudaMalloc: ") + cudaGetErrorString(e)); }
  cudaMemset(h->local, 0, sizeof(Block));
  cudaDeviceSynchronize();
  h->dev.rank = rank; h->dev.world = world;
  h->dev.blk[rank] = h->local;
  *out = h;
  return IEXA_OK;
}

// handles of MY x buffer (any cudaMalloc'ed pointer, e.g. a torch tensor; x_offset = its byte offset inside the
// allocation the handle names) and of my flag block — to be sent to every peer by the caller (torch.distributed)
int32_t iexa_halo_export(iexa_halo *h, const void *x_dev, unsigned char *handle_x, int64_t *x_offset, unsigned char *handle_flags) {
  if (!h || !x_dev || !handle_x || !x_offset || !handle_flags) HFAIL(IEXA_ERR_INVALID, "halo: null argument");
  HCK(cudaSetDevice(h->device));
  unsigned long long base = (unsigned long long)x_dev;
  static cuMemGetAddressRange_t get_range = [] {
    void *lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    return lib ? (cuMemGetAddressRange_t)dlsym(lib, "cuMemGetAddressRange_v2") : nullptr;
  }();
  size_t size = 0;
  if (get_range && get_range(&base, &size, (unsigned long long)x_dev) != 0) base = (unsigned long long)x_dev;
  cudaIpcMemHandle_t hx, hf;
  HCK(cudaIpcGetMemHandle(&hx, (void *)base));
  HCK(cudaIpcGetMemHandle(&hf, (void *)h->local));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  std::memcpy(handle_x, &hx, 64);
  std::memcpy(handle_flags, &hf, 64);
  *x_offset = (int64_t)((unsigned long long)x_dev - base);
  h->x_exported = x_dev;
  return IEXA_OK;
}

int32_t iexa_halo_connect(iexa_halo *h, int32_t peer, const unsigned char *handle_x, int64_t x_offset, const unsigned char *handle_flags) {
  if (!h || peer < 0 || peer >= h->world || peer == h->rank) HFAIL(IEXA_ERR_INVALID, "halo: bad peer");
  HCK(cudaSetDevice(h->device));
  cudaIpcMemHandle_t hx, hf;
  std::memcpy(&hx, handle_x, 64);
  std::memcpy(&hf, handle_flags, 64);
  void *px = nullptr, *pf = nullptr;
  HCK(cudaIpcOpenMemHandle(&px, hx, cudaIpcMemLazyEnablePeerAccess));
  h->opened.push_back(px);
  HCK(cudaIpcOpenMemHandle(&pf, hf, cudaIpcMemLazyEnablePeerAccess));
  h->opened.push_back(pf);
  h->dev.x[peer] = (double *)((char *)px + x_offset);
  h->dev.blk[peer] = (Block *)pf;
  return IEXA_OK;
}

// ranges [lo, hi) (0-based, global positions in x) this rank pushes to `peer` on every exchange
int32_t iexa_halo_set_sends(iexa_halo *h, int32_t peer, int64_t n_ranges, const int64_t *lo_hi) {
  if (!h || peer < 0 || peer >= h->world || peer == h->rank || n_ranges < 0 || (n_ranges > 0 && !lo_hi)) HFAIL(IEXA_ERR_INVALID, "halo: bad send table");
  std::vector<SendRange> keep;
  for (auto &s : h->sends) if (s.peer != peer) keep.push_back(s);
  for (int64_t i = 0; i < n_ranges; ++i)
    if (lo_hi[2 * i + 1] > lo_hi[2 * i]) keep.push_back(SendRange{peer, lo_hi[2 * i], lo_hi[2 * i + 1] - lo_hi[2 * i]});
  h->sends.swap(keep);
  h->tables_dirty = true;
  return IEXA_OK;
}
// the owners this rank receives from on every exchange
int32_t iexa_halo_set_recvs(iexa_halo *h, int32_t n_peers, const int32_t *peers) {
  if (!h || n_peers < 0 || n_peers > h->world) HFAIL(IEXA_ERR_INVALID, "halo: bad receive table");
  h->recv_peers.assign(peers, peers + n_peers);
  h->tables_dirty = true;
  return IEXA_OK;
}

static int32_t upload_tables(iexa_halo *h) {
  if (!h->tables_dirty) return IEXA_OK;
  h->send_peers.clear();
  for (auto &s : h->sends) {
    bool seen = false;
    for (int p : h->send_peers) seen = seen || p == s.peer;
    if (!seen) h->send_peers.push_back(s.peer);
  }
  for (int p : h->send_peers) if (!h->dev.blk[p] || !h->dev.x[p]) HFAIL(IEXA_ERR_STATE, "halo: send peer not connected");
  for (int p : h->recv_peers) if (p < 0 || p >= h->world || !h->dev.blk[p]) HFAIL(IEXA_ERR_STATE, "halo: receive peer not connected");
  if (h->send_peers.size() > 256 || h->recv_peers.size() > 256) HFAIL(IEXA_ERR_INVALID, "halo: too many peers");
  auto up = [&](void *&d, const void *src, size_t bytes) -> cudaError_t {
    if (d) cudaFree(d);
    d = nullptr;
    cudaError_t e = cudaMalloc(&d, bytes ? bytes : 8);
    if (e != cudaSuccess) return e;
    return bytes ? cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice) : cudaSuccess;
  };
  HCK(up(h->d_sends, h->sends.data(), h->sends.size() * sizeof(SendRange)));
  HCK(up(h->d_send_peers, h->send_peers.data(), h->send_peers.size() * sizeof(int)));
  HCK(up(h->d_recv_peers, h->recv_peers.data(), h->recv_peers.size() * sizeof(int)));
  h->dev.sends = (const SendRange *)h->d_sends; h->dev.nsend = (int)h->sends.size();
  h->dev.send_peers = (const int *)h->d_send_peers; h->dev.nsend_peers = (int)h->send_peers.size();
  h->dev.recv_peers = (const int *)h->d_recv_peers; h->dev.nrecv_peers = (int)h->recv_peers.size();
  h->tables_dirty = false;
  return IEXA_OK;
}

// collective: every rank calls it the same number of times.  x_dev must be the exported buffer.
int32_t iexa_halo_exchange(iexa_halo *h, const double *x_dev, void *stream) {
  if (!h || !x_dev) HFAIL(IEXA_ERR_INVALID, "halo: null argument");
  if (x_dev != h->x_exported) HFAIL(IEXA_ERR_INVALID, "halo: x is not the exported buffer (peers push into the exported one)");
  HCK(cudaSetDevice(h->device));
  int32_t rc = upload_tables(h);
  if (rc) return rc;
  ++h->epoch;
  if (h->dev.nsend_peers == 0 && h->dev.nrecv_peers == 0) return IEXA_OK;
  halo_exchange_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(h->dev, x_dev, h->epoch);
  HCK(cudaGetLastError());
  return IEXA_OK;
}

// buf[0..n) <- sum over all ranks (n <= 1024), in place, deterministic, asynchronous on `stream`; collective
int32_t iexa_halo_allreduce_small(iexa_halo *h, double *buf_dev, int32_t n, void *stream) {
  if (!h || !buf_dev || n < 0 || n > RED_MAX) HFAIL(IEXA_ERR_INVALID, "halo: small all-reduce takes at most 1024 doubles");
  HCK(cudaSetDevice(h->device));
  for (int p = 0; p < h->world; ++p) if (!h->dev.blk[p]) HFAIL(IEXA_ERR_STATE, "halo: all-reduce needs every peer connected");
  ++h->red_epoch;
  if (n == 0 || h->world == 1) return IEXA_OK;
  small_allreduce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(h->dev, buf_dev, n, h->red_epoch);
  HCK(cudaGetLastError());
  return IEXA_OK;
}

// 0 = fine; otherwise the code of the bounded wait that expired (1000+peer: ack, 2000+peer: halo, 3000+peer: all-reduce)
int64_t iexa_halo_status(iexa_halo *h) {
  if (!h) return -1;
  cudaSetDevice(h->device);
  unsigned long long st = 0;
  if (cudaMemcpy(&st, &h->local->status, 8, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return (int64_t)st;
}

int32_t iexa_halo_destroy(iexa_halo *h) {
  if (!h) return IEXA_OK;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (void *p : h->opened) cudaIpcCloseMemHandle(p);
  if (h->d_sends) cudaFree(h->d_sends);
  if (h->d_send_peers) cudaFree(h->d_send_peers);
  if (h->d_recv_peers) cudaFree(h->d_recv_peers);
  if (h->local) cudaFree(h->local);
  delete h;
  return IEXA_OK;
}

} // extern "C"

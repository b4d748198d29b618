// stream_probe.cu — how much HBM bandwidth does a B200 sustain when one kernel walks S read streams and
// W write streams at once (thread k touches element k of every stream: the access pattern of cons!, where every
// variable block / row block of the reference's x / c layout is its own stream)?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o stream_probe tools/stream_probe.cu && ./stream_probe
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

template <int S, int W>
__global__ void __launch_bounds__(128) probe(const double* __restrict__ in, double* __restrict__ out, long long n) {
  const long long k = (long long)blockIdx.x * 128 + threadIdx.x;
  if (k >= n) return;
  double v[S];
#pragma unroll
  for (int s = 0; s < S; ++s) v[s] = __ldg(in + (long long)s * n + k);
  double acc = 0.0;
#pragma unroll
  for (int s = 0; s < S; ++s) acc += v[s];
#pragma unroll
  for (int w = 0; w < W; ++w) out[(long long)w * n + k] = acc + w;
}

template <int S, int W>
void run(long long total_elems) {
  const long long n = total_elems / (S + W);
  double *in, *out;
  cudaMalloc(&in, (size_t)S * n * 8);
  cudaMalloc(&out, (size_t)W * n * 8);
  cudaMemset(in, 0, (size_t)S * n * 8);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  const int nb = (int)((n + 127) / 128);
  for (int i = 0; i < 3; ++i) probe<S, W><<<nb, 128>>>(in, out, n);
  cudaEventRecord(a);
  const int reps = 20;
  for (int i = 0; i < reps; ++i) probe<S, W><<<nb, 128>>>(in, out, n);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  ms /= reps;
  printf("S=%2d W=%2d  n=%lld  %.4f ms  %.0f GB/s\n", S, W, n, ms, (double)(S + W) * n * 8 / ms / 1e6);
  cudaFree(in); cudaFree(out);
}

// pure writes, the shape of the staged COO tiles: every warp writes `step` consecutive 256-byte rows (one tile of
// 32 supports x step slots), tiles of consecutive warps are adjacent
template <int STEP>
__global__ void __launch_bounds__(128) tile_writer(double* __restrict__ out, long long ntiles) {
  const long long tile = ((long long)blockIdx.x * 128 + threadIdx.x) >> 5;
  if (tile >= ntiles) return;
  double* dst = out + tile * (32ll * STEP) + (threadIdx.x & 31);
#pragma unroll
  for (int i = 0; i < STEP; ++i) dst[32 * i] = (double)i;
}
template <int STEP>
void run_tiles(long long total_elems) {
  const long long ntiles = total_elems / (32 * STEP);
  double* out;
  cudaMalloc(&out, (size_t)ntiles * 32 * STEP * 8);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  const int nb = (int)((ntiles * 32 + 127) / 128);
  for (int i = 0; i < 3; ++i) tile_writer<STEP><<<nb, 128>>>(out, ntiles);
  cudaEventRecord(a);
  const int reps = 20;
  for (int i = 0; i < reps; ++i) tile_writer<STEP><<<nb, 128>>>(out, ntiles);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  ms /= reps;
  printf("pure writes, tiles of 32 x %2d doubles: %.4f ms  %.0f GB/s\n", STEP, ms, (double)ntiles * 32 * STEP * 8 / ms / 1e6);
  cudaFree(out);
}

// the shape of jac_coord!: every thread reads S streams at its support, every warp writes one tile of 32 x STEP doubles
template <int S, int STEP>
__global__ void __launch_bounds__(128) mix(const double* __restrict__ in, double* __restrict__ out, long long n) {
  const long long k = (long long)blockIdx.x * 128 + threadIdx.x;
  if (k >= n) return;
  double acc = 0.0;
#pragma unroll
  for (int s = 0; s < S; ++s) acc += __ldg(in + (long long)s * n + k);
  double* dst = out + (k >> 5) * (32ll * STEP) + (threadIdx.x & 31);
#pragma unroll
  for (int i = 0; i < STEP; ++i) dst[32 * i] = acc + i;
}
template <int S, int STEP>
void run_mix(long long n) {
  double *in, *out;
  cudaMalloc(&in, (size_t)S * n * 8); cudaMalloc(&out, (size_t)STEP * n * 8);
  cudaMemset(in, 0, (size_t)S * n * 8);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int nb = (int)((n + 127) / 128);
  for (int i = 0; i < 3; ++i) mix<S, STEP><<<nb, 128>>>(in, out, n);
  cudaEventRecord(a);
  for (int i = 0; i < 20; ++i) mix<S, STEP><<<nb, 128>>>(in, out, n);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); ms /= 20;
  printf("mix: %d read streams + tiles of 32 x %d doubles, n=%lld: %.4f ms  %.0f GB/s\n", S, STEP, n, ms, (double)(S + STEP) * n * 8 / ms / 1e6);
  cudaFree(in); cudaFree(out);
}

int main() {
  run_mix<8, 35>(4000000); run_mix<8, 72>(2000000); run_mix<4, 35>(4000000); run_mix<21, 59>(2000000); run_mix<19, 9>(6000000);
  run_tiles<1>(150000000); run_tiles<10>(150000000); run_tiles<35>(150000000);
  {
    double* p; cudaMalloc(&p, 1200000000);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) cudaMemsetAsync(p, 0, 1200000000);
    cudaEventRecord(a);
    for (int i = 0; i < 20; ++i) cudaMemsetAsync(p, 0, 1200000000);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("cudaMemset 1.2 GB: %.4f ms  %.0f GB/s\n", ms / 20, 1200000000.0 / (ms / 20) / 1e6);
    cudaFree(p);
  }
  const long long E = 180000000; // 1.44 GB of traffic per launch
  run<1, 1>(E); run<2, 2>(E); run<4, 4>(E); run<8, 8>(E); run<16, 16>(E); run<20, 20>(E); run<32, 16>(E);
  run<20, 9>(E); run<8, 1>(E); run<16, 1>(E); run<32, 1>(E); run<1, 8>(E); run<1, 16>(E); run<1, 32>(E); run<4, 36>(E);
  return 0;
}

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

int main() {
  const long long E = 180000000; // 1.44 GB of traffic per launch
  run<1, 1>(E); run<2, 2>(E); run<4, 4>(E); run<8, 8>(E); run<16, 16>(E); run<20, 20>(E); run<32, 16>(E);
  run<20, 9>(E); run<8, 1>(E); run<16, 1>(E); run<32, 1>(E); run<1, 8>(E); run<1, 16>(E); run<1, 32>(E); run<4, 36>(E);
  return 0;
}

// Dependent-chain latency of FP64 / FP32 / DMMA instructions on one warp (clock64 around an unrolled chain).
#include <cstdio>
#include <cuda_runtime.h>
template <int N> __global__ void k_dfma(double* out, double a, double b, long long* cyc) {
  double x = out[threadIdx.x];
  long long t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) x = fma(x, a, b);
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int N> __global__ void k_ffma(float* out, float a, float b, long long* cyc) {
  float x = out[threadIdx.x];
  long long t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) x = fmaf(x, a, b);
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int N> __global__ void k_dmma(double* out, double a, double b, long long* cyc) {
  double c0 = out[threadIdx.x], c1 = 0.0;
  long long t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
  long long t1 = clock64();
  out[threadIdx.x] = c0 + c1; if (threadIdx.x == 0) *cyc = t1 - t0;
}
// ILP variant: K independent chains per thread
template <int N, int K> __global__ void k_dfma_ilp(double* out, double a, double b, long long* cyc) {
  double x[K];
  for (int j = 0; j < K; ++j) x[j] = out[threadIdx.x] + j;
  long long t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < K; ++j) x[j] = fma(x[j], a, b);
  long long t1 = clock64();
  double s = 0; for (int j = 0; j < K; ++j) s += x[j];
  out[threadIdx.x] = s; if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double* d; float* f; long long* c; long long h;
  cudaMalloc(&d, 4096 * 8); cudaMalloc(&f, 4096 * 4); cudaMalloc(&c, 8);
  cudaMemset(d, 0, 4096 * 8); cudaMemset(f, 0, 4096 * 4);
  constexpr int N = 256;
  for (int warps : {1, 2, 4, 8}) {
    k_dfma<N><<<1, 32 * warps>>>(d, 1.0000001, 1e-9, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("DFMA dependent chain, %d warp(s)/CTA: %.1f cycles per op\n", warps, (double)h / N);
  }
  k_ffma<N><<<1, 32>>>(f, 1.0000001f, 1e-9f, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("FFMA dependent chain: %.1f cycles per op\n", (double)h / N);
  for (int warps : {1, 4, 8}) {
    k_dmma<N><<<1, 32 * warps>>>(d, 1.0000001, 1e-9, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("DMMA m8n8k4 dependent chain, %d warp(s): %.1f cycles per op\n", warps, (double)h / N);
  }
  k_dfma_ilp<64, 2><<<1, 32>>>(d, 1.0000001, 1e-9, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("DFMA 2 chains/thread, 1 warp: %.1f cycles per op\n", (double)h / (64 * 2));
  k_dfma_ilp<64, 8><<<1, 32>>>(d, 1.0000001, 1e-9, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("DFMA 8 chains/thread, 1 warp: %.1f cycles per op\n", (double)h / (64 * 8));
  k_dfma_ilp<32, 32><<<1, 32>>>(d, 1.0000001, 1e-9, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("DFMA 32 chains/thread, 1 warp: %.1f cycles per op\n", (double)h / (32 * 32));
  k_dfma_ilp<32, 32><<<1, 256>>>(d, 1.0000001, 1e-9, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("DFMA 32 chains/thread, 8 warps: %.1f cycles per op per warp\n", (double)h / (32 * 32));
  return 0;
}

// Pipe-ceiling microbenchmark for sm_100a: FP64 FMA (DFMA), FP64 tensor (DMMA, all
// mma.sync f64 shapes), FP32 FMA (FFMA).  MEASURED_PEAKS.json only carries HBM and bf16
// figures; the expm / Frechet kernels are FP64- or FP32-pipe bound, so
// their roofline denominators are measured here, on the same box, and written as JSON.
//
// build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp_peaks tools/fp_peaks.cu
// run:    tools/fp_peaks > gpurun_out/fp_peaks.json
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;

// 16 independent DFMA chains per thread.
__global__ void __launch_bounds__(256) k_dfma(double* out, double a, double b) {
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) k_ffma(float* out, float a, float b) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-6f + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  if (s == 123.456f) out[0] = s;
}

// DMMA m8n8k4: 8 independent accumulator tiles per warp.
__global__ void __launch_bounds__(256) k_dmma884(double* out, double a, double b) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = threadIdx.x; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}

// DMMA m16n8k4: A 2 regs, B 1 reg, C 4 regs.
__global__ void __launch_bounds__(256) k_dmma1684(double* out, double a, double b) {
  double c[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = threadIdx.x; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a), "d"(b), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456) out[0] = s;
}

// DMMA m16n8k8: A 4 regs, B 2 regs, C 4 regs.
__global__ void __launch_bounds__(256) k_dmma1688(double* out, double a, double b) {
  double c[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = threadIdx.x; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456) out[0] = s;
}

// DMMA m16n8k16: A 8 regs, B 4 regs, C 4 regs.
__global__ void __launch_bounds__(256) k_dmma16816(double* out, double a, double b) {
  double c[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = threadIdx.x; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, "
                   "{%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b),
                     "d"(b), "d"(a), "d"(b), "d"(a));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456) out[0] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) launch();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
  CK(cudaGetLastError());
  return ms / reps;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  double* dout; float* fout;
  CK(cudaMalloc(&dout, 64)); CK(cudaMalloc(&fout, 64));
  printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
  const int reps = 20;
  for (int bps = 1; bps <= 8; bps *= 2) {       // blocks per SM (256 threads each)
    int grid = sms * bps;
    double warps = (double)grid * 8;
    double ms;
    ms = time_ms([&] { k_dfma<<<grid, 256>>>(dout, 1.0000001, 1e-9); }, reps);
    printf(",\n \"dfma_tflops_bps%d\": %.3f", bps, warps * 32 * 16.0 * ITERS * 2 / ms / 1e9);
    ms = time_ms([&] { k_ffma<<<grid, 256>>>(fout, 1.0000001f, 1e-9f); }, reps);
    printf(", \"ffma_tflops_bps%d\": %.3f", bps, warps * 32 * 16.0 * ITERS * 2 / ms / 1e9);
    ms = time_ms([&] { k_dmma884<<<grid, 256>>>(dout, 1.0000001, 1e-9); }, reps);
    printf(", \"dmma_m8n8k4_tflops_bps%d\": %.3f", bps, warps * 8.0 * ITERS * (8 * 8 * 4) * 2 / ms / 1e9);
    ms = time_ms([&] { k_dmma1684<<<grid, 256>>>(dout, 1.0000001, 1e-9); }, reps);
    printf(", \"dmma_m16n8k4_tflops_bps%d\": %.3f", bps, warps * 8.0 * ITERS * (16 * 8 * 4) * 2 / ms / 1e9);
    ms = time_ms([&] { k_dmma1688<<<grid, 256>>>(dout, 1.0000001, 1e-9); }, reps);
    printf(", \"dmma_m16n8k8_tflops_bps%d\": %.3f", bps, warps * 8.0 * ITERS * (16 * 8 * 8) * 2 / ms / 1e9);
    ms = time_ms([&] { k_dmma16816<<<grid, 256>>>(dout, 1.0000001, 1e-9); }, reps);
    printf(", \"dmma_m16n8k16_tflops_bps%d\": %.3f", bps, warps * 8.0 * ITERS * (16 * 8 * 16) * 2 / ms / 1e9);
  }
  // sustained: ~3 s of back-to-back DFMA and DMMA launches (power-capped clocks)
  {
    int grid = sms * 4;
    double warps = (double)grid * 8;
    double one = time_ms([&] { k_dfma<<<grid, 256>>>(dout, 1.0000001, 1e-9); }, 5);
    int n = (int)(3000.0 / one) + 1;
    double ms = time_ms([&] { k_dfma<<<grid, 256>>>(dout, 1.0000001, 1e-9); }, n);
    printf(",\n \"dfma_tflops_sustained\": %.3f", warps * 32 * 16.0 * ITERS * 2 / ms / 1e9);
    one = time_ms([&] { k_dmma884<<<grid, 256>>>(dout, 1.0000001, 1e-9); }, 5);
    n = (int)(3000.0 / one) + 1;
    ms = time_ms([&] { k_dmma884<<<grid, 256>>>(dout, 1.0000001, 1e-9); }, n);
    printf(", \"dmma_m8n8k4_tflops_sustained\": %.3f", warps * 8.0 * ITERS * 256 * 2 / ms / 1e9);
    one = time_ms([&] { k_ffma<<<grid, 256>>>(fout, 1.0000001f, 1e-9f); }, 5);
    n = (int)(3000.0 / one) + 1;
    ms = time_ms([&] { k_ffma<<<grid, 256>>>(fout, 1.0000001f, 1e-9f); }, n);
    printf(", \"ffma_tflops_sustained\": %.3f", warps * 32 * 16.0 * ITERS * 2 / ms / 1e9);
  }
  printf("\n}\n");
  return 0;
}

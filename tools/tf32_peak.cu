// Peak of the legacy tensor path (mma.sync m16n8k8 TF32, FP32 accumulate) on this GPU: independent accumulator
// chains per warp, no memory traffic.  Used to decide whether an error-compensated 3xTF32 complex64 GEMM can beat
// the FP32 FMA pipe (71.8 TFLOP/s measured).
#include <cstdio>
#include <cuda_runtime.h>
template <int CHAINS>
__global__ void k(float* out, int iters) {
  float c[CHAINS][4];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
  unsigned a0 = threadIdx.x, a1 = threadIdx.x + 1, a2 = 3, a3 = 4, b0 = 5, b1 = 6;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int bps : {1, 2, 4}) for (int threads : {128, 256, 512}) {
    k<8><<<148 * bps, threads>>>(out, 100);
    cudaEventRecord(e0); k<8><<<148 * bps, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 16 * 8 * 8 * 8.0 * iters * (threads / 32) * 148.0 * bps;
    printf("mma.sync m16n8k8 tf32: %d CTA/SM x %d threads: %.1f TFLOP/s\n", bps, threads, flops / ms / 1e9);
  }
  return 0;
}

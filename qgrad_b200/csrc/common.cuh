// Shared device/host helpers for the qgrad_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include "../../include/qgrad_b200.h"

namespace qg {

// ------------------------------------------------------------------ complex scalar
template <typename T> struct Vec2;
template <> struct Vec2<float>  { using type = float2; };
template <> struct Vec2<double> { using type = double2; };

template <typename T>
struct alignas(2 * sizeof(T)) cx {
  T re, im;
  __host__ __device__ cx() {}
  __host__ __device__ cx(T r, T i) : re(r), im(i) {}
};

template <typename T> __host__ __device__ __forceinline__ cx<T> operator+(cx<T> a, cx<T> b) { return cx<T>(a.re + b.re, a.im + b.im); }
template <typename T> __host__ __device__ __forceinline__ cx<T> operator-(cx<T> a, cx<T> b) { return cx<T>(a.re - b.re, a.im - b.im); }
template <typename T> __host__ __device__ __forceinline__ cx<T> operator-(cx<T> a) { return cx<T>(-a.re, -a.im); }
template <typename T> __host__ __device__ __forceinline__ cx<T> operator*(cx<T> a, cx<T> b) {
  return cx<T>(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
template <typename T> __host__ __device__ __forceinline__ cx<T> operator*(T s, cx<T> a) { return cx<T>(s * a.re, s * a.im); }
template <typename T> __host__ __device__ __forceinline__ cx<T> conj(cx<T> a) { return cx<T>(a.re, -a.im); }
// acc += a*b  (4 FMAs)
template <typename T> __device__ __forceinline__ void cfma(cx<T>& acc, cx<T> a, cx<T> b) {
  acc.re = fma(a.re, b.re, acc.re); acc.re = fma(-a.im, b.im, acc.re);
  acc.im = fma(a.re, b.im, acc.im); acc.im = fma(a.im, b.re, acc.im);
}
// acc += conj(a)*b
template <typename T> __device__ __forceinline__ void cfma_conj(cx<T>& acc, cx<T> a, cx<T> b) {
  acc.re = fma(a.re, b.re, acc.re); acc.re = fma(a.im, b.im, acc.re);
  acc.im = fma(a.re, b.im, acc.im); acc.im = fma(-a.im, b.re, acc.im);
}
template <typename T> __device__ __forceinline__ cx<T> crecip(cx<T> a) {
  T d = T(1) / (a.re * a.re + a.im * a.im);
  return cx<T>(a.re * d, -a.im * d);
}
__device__ __forceinline__ float  cabs_(cx<float> a)  { return hypotf(a.re, a.im); }
__device__ __forceinline__ double cabs_(cx<double> a) { return hypot(a.re, a.im); }
__device__ __forceinline__ void sincos_(float x, float* s, float* c)   { sincosf(x, s, c); }
__device__ __forceinline__ void sincos_(double x, double* s, double* c) { sincos(x, s, c); }

template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <typename T> __device__ __forceinline__ T warp_max(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------ Pade plan
// Orders 3/5/7 only (complex128) and 3/5 (complex64): every bound is < 2 ln 2, so
// ||q_m(A)/b_0 - I||_1 <= e^{||A||_1/2} - 1 < 1, i.e. the Pade denominator is strictly column
// diagonally dominant.  Partial pivoting then provably selects the diagonal at every step of the
// elimination (dominance is inherited by Schur complements), so the solve runs unpivoted with
// growth factor <= 2.  See DESIGN.md "Pade plan".
template <typename T> struct PadePlan;
template <> struct PadePlan<double> {
  static constexpr int kTop = 7;
  // value-only bounds theta_m (Higham 2005) and value+Frechet bounds ell_m (Al-Mohy & Higham 2009)
  __host__ __device__ static double bound(int m, bool frechet) {
    return frechet ? (m == 3 ? 1.08e-2 : m == 5 ? 2.00e-1 : 7.83e-1)
                   : (m == 3 ? 1.495585217958292e-2 : m == 5 ? 2.539398330063230e-1 : 9.504178996162932e-1);
  }
};
template <> struct PadePlan<float> {
  static constexpr int kTop = 5;
  __host__ __device__ static double bound(int m, bool frechet) {
    return frechet ? (m == 3 ? 3.4e-1 : 1.3) : (m == 3 ? 4.258730016922831e-1 : 1.3);
  }
};

// b_k / b_0 of the [m/m] Pade numerator: coefficient k of the normalised polynomial.
__host__ __device__ inline double pade_coef(int m, int k) {
  // m = 3: 120,60,12,1   m = 5: 30240,15120,3360,420,30,1
  // m = 7: 17297280,8648640,1995840,277200,25200,1512,56,1
  if (k > m) return 0.0;
  if (m == 3) { const double b[4] = {1.0, 0.5, 0.1, 1.0 / 120.0}; return b[k]; }
  if (m == 5) { const double b[6] = {1.0, 0.5, 1.0 / 9.0, 1.0 / 72.0, 1.0 / 1008.0, 1.0 / 30240.0}; return b[k]; }
  const double b[8] = {1.0, 0.5, 1995840.0 / 17297280.0, 277200.0 / 17297280.0, 25200.0 / 17297280.0,
                       1512.0 / 17297280.0, 56.0 / 17297280.0, 1.0 / 17297280.0};
  return b[k];
}

// (m, s) from the 1-norm.  s < 0 never returned; s > max_sq is reported by the caller as NaN.
template <typename T>
__host__ __device__ inline void pade_plan(double norm1, bool frechet, int* m, int* s) {
  const int top = PadePlan<T>::kTop;
  for (int mm = 3; mm <= top; mm += 2) {
    if (norm1 <= PadePlan<T>::bound(mm, frechet)) { *m = mm; *s = 0; return; }
  }
  const double b = PadePlan<T>::bound(top, frechet);
  int ss = 0;
  double x = norm1;
  // ceil(log2(norm/b)) without libm edge cases; bail out at 64 (inf/NaN norms land here too)
  while (!(x <= b) && ss < 64) { x *= 0.5; ++ss; }
  *m = top; *s = ss;
}

// ------------------------------------------------------------------ host-side plumbing
extern std::atomic<int64_t> g_launch_count;
inline void count_launch(int n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }

#define QG_RETURN_IF_CUDA(expr)                         \
  do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return (int)e__; } while (0)
#define QG_CHECK_LAUNCH()                               \
  do { qg::count_launch(); cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return (int)e__; } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace qg

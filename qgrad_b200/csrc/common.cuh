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
// Next pivot's reciprocal 1 / (dn - w * pinv) -- THE sequential chain of the Gauss-Jordan sweep (one per
// step, and FP64 results take ~100 cycles to come back on this part).  FP64: the pivot by two chained
// FMAs per component; |p|^2; then ONE cubic (Halley-type) correction r0 (1 + e + e^2), e = 1 - |p|^2 r0,
// of a seed r0 ~ 1/|p|^2 that a float32 shadow of the same expression has ready long before:
// 2^-22 -> 2^-66.  Eight dependent FP64 results instead of the twelve of "multiply, subtract, divide".
__device__ __forceinline__ cx<float> next_pivot_recip(cx<float> dn, cx<float> w, cx<float> pinv) {
  const float pre = fmaf(w.im, pinv.im, fmaf(-w.re, pinv.re, dn.re));
  const float pim = fmaf(-w.im, pinv.re, fmaf(-w.re, pinv.im, dn.im));
  const float r = __frcp_rn(fmaf(pre, pre, pim * pim));
  return cx<float>(pre * r, -pim * r);
}
__device__ __forceinline__ cx<double> next_pivot_recip(cx<double> dn, cx<double> w, cx<double> pinv) {
  const float fre = fmaf((float)w.im, (float)pinv.im, fmaf(-(float)w.re, (float)pinv.re, (float)dn.re));
  const float fim = fmaf(-(float)w.im, (float)pinv.re, fmaf(-(float)w.re, (float)pinv.im, (float)dn.im));
  const double r0 = (double)__frcp_rn(fmaf(fre, fre, fim * fim));
  const double pre = fma(w.im, pinv.im, fma(-w.re, pinv.re, dn.re));
  const double pim = fma(-w.im, pinv.re, fma(-w.re, pinv.im, dn.im));
  const double d = fma(pre, pre, pim * pim);
  const double e = fma(-d, r0, 1.0);
  const double r = fma(r0, fma(e, e, e), r0);
  return cx<double>(pre * r, -pim * r);
}
template <typename T> __device__ __forceinline__ cx<T> crecip_pivot(cx<T> a) {
  return next_pivot_recip(a, cx<T>(T(0), T(0)), cx<T>(T(0), T(0)));
}
__device__ __forceinline__ float  cabs_(cx<float> a)  { return hypotf(a.re, a.im); }
__device__ __forceinline__ double cabs_(cx<double> a) { return hypot(a.re, a.im); }
// |a| for the NORMS of the expm kernels: sqrt(re^2 + im^2) without hypot's range scaling (a third of its FP64
// instructions, on the pipe the products need).  Squares overflow only for |a| > 1e154 (1e19 in float); the norm is
// then inf and the matrix is flagged as needing too many squarings, which is where such an input ends up anyway.
__device__ __forceinline__ float  cabs_norm(cx<float> a)  { return sqrtf(fmaf(a.re, a.re, a.im * a.im)); }
__device__ __forceinline__ double cabs_norm(cx<double> a) { return sqrt(fma(a.re, a.re, a.im * a.im)); }
template <typename T> __device__ __forceinline__ T cabs2_(cx<T> a) { return fma(a.re, a.re, a.im * a.im); }
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

// ------------------------------------------------------------------ in-place Gauss-Jordan inverse in shared memory
// Unpivoted (the Pade denominator needs no pivoting, see below), NP x NP.  (NP/B)^2 threads take part, each
// keeping a B x B block of the matrix in REGISTERS for the whole sweep; per step only the pivot row,
// the pivot column and two scalars travel through shared memory, double-buffered by step parity, so a
// step costs ONE barrier and, per thread, B complex multiplies + B*B complex FMAs.  The step loop is
// unrolled by B so every register index is static.  A single thread (one outside the worker set when
// there is one) forms the NEXT pivot a_{k+1,k+1} - a_{k+1,k} a_{k,k+1} / a_kk from the broadcast
// vectors and takes its reciprocal while the others update: the division is off the critical path.
//   idx(r, c): element offset of (r, c) in M.   scratch: 4 NP + 4 complex values.
//   sync(): barrier over ALL NT threads of the CTA (every thread must call this function).
template <typename T, int NP, int B, int NT, typename Idx, typename Sync>
__device__ __forceinline__ void gj_inverse_smem(cx<T>* M, cx<T>* scratch, Idx idx, Sync sync) {
  constexpr int NB = NP / B, NWORK = NB * NB;
  static_assert(NB * B == NP && NWORK <= NT, "block split");
  const int tid = threadIdx.x;
  const bool worker = tid < NWORK;
  const int bi = tid / NB, bj = tid % NB;
  const int recip_tid = NWORK < NT ? NWORK : 0;
  cx<T>* rowb = scratch;            // [2][B][NB]: column c = bj B + b at b NB + bj, so a warp's 16-byte loads are contiguous
  cx<T>* colb = scratch + 2 * NP;   // [2][NP]
  cx<T>* dnext = scratch + 4 * NP;  // [2]
  cx<T>* pinvb = scratch + 4 * NP + 2;
  cx<T> v[B][B];
  if (worker) {
#pragma unroll
    for (int a = 0; a < B; ++a)
#pragma unroll
      for (int b = 0; b < B; ++b) v[a][b] = M[idx(bi * B + a, bj * B + b)];
    if (bi == 0) {
#pragma unroll
      for (int b = 0; b < B; ++b) rowb[b * NB + bj] = v[0][b];
    }
    if (bj == 0) {
#pragma unroll
      for (int a = 0; a < B; ++a) colb[bi * B + a] = v[a][0];
    }
    if (bi == 0 && bj == 0) { pinvb[0] = crecip_pivot(v[0][0]); dnext[0] = v[B > 1 ? 1 : 0][B > 1 ? 1 : 0]; }
  }
  sync();
  for (int kk = 0; kk < NB; ++kk) {
#pragma unroll
    for (int kq = 0; kq < B; ++kq) {
      const int k = kk * B + kq, par = k & 1;
      const cx<T> pinv = pinvb[par];
      if (tid == recip_tid && k + 1 < NP) {
        const cx<T> w = colb[par * NP + k + 1] * rowb[par * NP + ((k + 1) % B) * NB + (k + 1) / B];   // off the pinv chain
        pinvb[par ^ 1] = next_pivot_recip(dnext[par], w, pinv);
      }
      if (worker) {
        cx<T> t[B], h[B];
#pragma unroll
        for (int b = 0; b < B; ++b) t[b] = rowb[par * NP + b * NB + bj] * pinv;
#pragma unroll
        for (int a = 0; a < B; ++a) h[a] = colb[par * NP + bi * B + a];
        // row k: a'_kc = a_kc / p (a'_kk = 1/p);  column k: a'_rk = -a_rk / p;  else a'_rc = a_rc - a_rk a_kc / p
        if (bj == kk) {
          t[kq] = pinv;
#pragma unroll
          for (int a = 0; a < B; ++a) v[a][kq] = cx<T>(T(0), T(0));
        }
        if (bi == kk) {
          h[kq] = cx<T>(T(-1), T(0));
#pragma unroll
          for (int b = 0; b < B; ++b) v[kq][b] = cx<T>(T(0), T(0));
        }
#pragma unroll
        for (int a = 0; a < B; ++a) {
          const cx<T> nh = -h[a];
#pragma unroll
          for (int b = 0; b < B; ++b) cfma(v[a][b], nh, t[b]);
        }
        // broadcast row / column k+1 and the diagonal element k+2 for the next steps
        constexpr int dummy = 0; (void)dummy;
        const int k1q = (kq + 1) % B, k1b = kk + (kq + 1) / B;
        const int k2q = (kq + 2) % B, k2b = kk + (kq + 2) / B;
        if (bi == k1b) {
#pragma unroll
          for (int b = 0; b < B; ++b) rowb[(par ^ 1) * NP + b * NB + bj] = v[k1q][b];
        }
        if (bj == k1b) {
#pragma unroll
          for (int a = 0; a < B; ++a) colb[(par ^ 1) * NP + bi * B + a] = v[a][k1q];
        }
        if (bi == k2b && bj == k2b) dnext[par ^ 1] = v[k2q][k2q];
      }
      sync();
    }
  }
  if (worker) {
#pragma unroll
    for (int a = 0; a < B; ++a)
#pragma unroll
      for (int b = 0; b < B; ++b) M[idx(bi * B + a, bj * B + b)] = v[a][b];
  }
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

// "set this kernel attribute once per DEVICE": true the first time it is called on the current device for `done`
inline bool first_use_on_device(std::atomic<unsigned long long>& done) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return true;
  const unsigned long long bit = 1ull << (dev & 63);
  if (done.load(std::memory_order_relaxed) & bit) return false;
  done.fetch_or(bit, std::memory_order_relaxed);
  return true;
}

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace qg

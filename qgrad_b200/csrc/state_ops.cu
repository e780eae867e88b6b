// K4 -- batched expectation values and fidelities, forward and reverse mode.
// Reference semantics: expect/_expect_ket/_expect_dm (qgrad/qgrad_qutip.py:166-203) and
// fidelity/_fidelity_ket/_fidelity_dm (qgrad/qgrad_qutip.py:11-64).
//
// All of these move O(n^2) or O(n) bytes per O(n^2) / O(n) flops: HBM-bound.  One sample is owned
// by a sub-warp group of TPS = min(32, next_pow2(n)) lanes (n = 2 -> 16 samples per warp), rows
// strided over the lanes, group reductions by shuffle.  No fast-math: products of exact 0/1
// inputs must stay exact (the reference tests assert == 0.0 / == 1.0 on basis states).
#include "common.cuh"
#include <cstdint>

namespace qg {
namespace state {

__host__ __device__ inline int group_size(int n) { int t = 1; while (t < n && t < 32) t <<= 1; return t; }

template <typename T> __device__ __forceinline__ T group_sum(T v, int tps) {
  for (int o = tps >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// out[b] = sum_i conj(psi_i) sum_j O_ij psi_j
template <typename T>
__global__ void expect_ket_fwd_kernel(const cx<T>* __restrict__ O, const cx<T>* __restrict__ psi, cx<T>* __restrict__ out,
                                      long long batch, int n, int ob, int tps) {
  const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long b = gt / tps;
  const int l = (int)(gt % tps);
  const bool live = b < batch;
  cx<T> acc(T(0), T(0));
  if (live) {
    const cx<T>* o = O + (ob ? (size_t)b * n * n : 0);
    const cx<T>* p = psi + (size_t)b * n;
    if (ob && n >= 16) {
      // per-sample operator: stream it in memory order (coalesced), the ket comes from cache
      int i = l / n, j = l - i * n;
      const int di = tps / n, dj = tps - di * n;                   // e += tps without a division per element
#pragma unroll 4
      for (int e = l; e < n * n; e += tps) {
        cx<T> t(T(0), T(0));
        cfma(t, o[e], p[j]);
        cfma_conj(acc, p[i], t);
        i += di; j += dj;
        if (j >= n) { j -= n; ++i; }
      }
    } else {
      for (int i = l; i < n; i += tps) {
        cx<T> row(T(0), T(0));
        for (int j = 0; j < n; ++j) cfma(row, o[(size_t)i * n + j], p[j]);
        cfma_conj(acc, p[i], row);
      }
    }
  }
  acc.re = group_sum(acc.re, tps); acc.im = group_sum(acc.im, tps);
  if (live && l == 0) out[b] = acc;
}

// ketbar_i = conj(g) (O psi)_i + g (O^H psi)_i ;  operbar_ij (+)= g psi_i conj(psi_j)
template <typename T>
__global__ void expect_ket_bwd_kernel(const cx<T>* __restrict__ O, const cx<T>* __restrict__ psi,
                                      const cx<T>* __restrict__ gbar, cx<T>* __restrict__ ketbar,
                                      cx<T>* __restrict__ operbar, long long batch, int n, int ob, int tps) {
  const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long b = gt / tps;
  const int l = (int)(gt % tps);
  if (b >= batch) return;
  const cx<T>* o = O + (ob ? (size_t)b * n * n : 0);
  const cx<T>* p = psi + (size_t)b * n;
  const cx<T> g = gbar[b];
  for (int i = l; i < n; i += tps) {
    if (ketbar) {
      cx<T> r1(T(0), T(0)), r2(T(0), T(0));
      for (int j = 0; j < n; ++j) { cfma(r1, o[(size_t)i * n + j], p[j]); cfma_conj(r2, o[(size_t)j * n + i], p[j]); }
      ketbar[(size_t)b * n + i] = conj(g) * r1 + g * r2;
    }
    if (operbar) {
      const cx<T> gp = g * p[i];
      for (int j = 0; j < n; ++j) {
        const cx<T> v = gp * conj(p[j]);
        if (ob) operbar[(size_t)b * n * n + (size_t)i * n + j] = v;
        else { atomicAdd(&operbar[(size_t)i * n + j].re, v.re); atomicAdd(&operbar[(size_t)i * n + j].im, v.im); }
      }
    }
  }
}

// out[b] = sum_ij rho_ij O_ji.  n >= 16: the group streams rho in memory order (coalesced) against O^T,
// which a shared operator provides from shared memory (staged transposed once per block: sOt, else null)
// and a per-sample operator from global memory (every fetched sector is used within the sample).
// n < 16: rows strided over the lanes (whole rows fit a few sectors).
template <typename T>
__global__ void expect_dm_fwd_kernel(const cx<T>* __restrict__ O, const cx<T>* __restrict__ rho, cx<T>* __restrict__ out,
                                     long long batch, int n, int ob, int tps, int stage) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<T>* sOt = reinterpret_cast<cx<T>*>(smem_raw);
  if (stage) {
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) { const int i = e / n, j = e - i * n; sOt[e] = O[(size_t)j * n + i]; }
    __syncthreads();
  }
  const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long b = gt / tps;
  const int l = (int)(gt % tps);
  const bool live = b < batch;
  cx<T> acc(T(0), T(0));
  if (live) {
    const cx<T>* o = O + (ob ? (size_t)b * n * n : 0);
    const cx<T>* r = rho + (size_t)b * n * n;
    if (stage) {
      for (int e = l; e < n * n; e += tps) cfma(acc, r[e], sOt[e]);
    } else if (n >= 16) {
      for (int e = l; e < n * n; e += tps) { const int i = e / n, j = e - i * n; cfma(acc, r[e], o[(size_t)j * n + i]); }
    } else {
      for (int i = l; i < n; i += tps)
        for (int j = 0; j < n; ++j) cfma(acc, r[(size_t)i * n + j], o[(size_t)j * n + i]);
    }
  }
  acc.re = group_sum(acc.re, tps); acc.im = group_sum(acc.im, tps);
  if (live && l == 0) out[b] = acc;
}

// rhobar_ij = g conj(O_ji) ; operbar_ji (+)= g conj(rho_ij).  n >= 16: each output is written in ITS memory
// order (coalesced stores), the transposed operand coming from shared memory (shared operator) or cache.
template <typename T>
__global__ void expect_dm_bwd_kernel(const cx<T>* __restrict__ O, const cx<T>* __restrict__ rho,
                                     const cx<T>* __restrict__ gbar, cx<T>* __restrict__ rhobar,
                                     cx<T>* __restrict__ operbar, long long batch, int n, int ob, int tps, int stage) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<T>* sOt = reinterpret_cast<cx<T>*>(smem_raw);
  if (stage) {
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) { const int i = e / n, j = e - i * n; sOt[e] = O[(size_t)j * n + i]; }
    __syncthreads();
  }
  const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long b = gt / tps;
  const int l = (int)(gt % tps);
  if (b >= batch) return;
  const cx<T>* o = O + (ob ? (size_t)b * n * n : 0);
  const cx<T>* r = rho + (size_t)b * n * n;
  const cx<T> g = gbar[b];
  if (n >= 16) {
    for (int e = l; e < n * n; e += tps) {
      const int i = e / n, j = e - i * n;
      if (rhobar) rhobar[(size_t)b * n * n + e] = g * conj(stage ? sOt[e] : o[(size_t)j * n + i]);
      if (operbar) {
        const cx<T> v = g * conj(r[(size_t)j * n + i]);            // operbar_(i,j) = g conj(rho_(j,i))
        if (ob) operbar[(size_t)b * n * n + e] = v;
        else { atomicAdd(&operbar[e].re, v.re); atomicAdd(&operbar[e].im, v.im); }
      }
    }
    return;
  }
  for (int i = l; i < n; i += tps)
    for (int j = 0; j < n; ++j) {
      if (rhobar) rhobar[(size_t)b * n * n + (size_t)i * n + j] = g * conj(o[(size_t)j * n + i]);
      if (operbar) {
        const cx<T> v = g * conj(r[(size_t)i * n + j]);
        if (ob) operbar[(size_t)b * n * n + (size_t)j * n + i] = v;
        else { atomicAdd(&operbar[(size_t)j * n + i].re, v.re); atomicAdd(&operbar[(size_t)j * n + i].im, v.im); }
      }
    }
}

// shared operator, rhobar only, n < 16: rhobar[b] = g[b] conj(O^T) is a broadcast store.  A thread owns ONE position
// e of the matrix (its conj(O_ji) stays in a register) and walks kBcastMats samples; 256 / span samples sit side by
// side in a block so that a warp's stores are contiguous.
constexpr int kBcastMats = 64;
template <typename T>
__global__ void expect_dm_bwd_bcast_kernel(const cx<T>* __restrict__ O, const cx<T>* __restrict__ gbar,
                                           cx<T>* __restrict__ rhobar, long long batch, int n) {
  const int nn = n * n;
  int span = 1;
  while (span < nn) span <<= 1;                                    // nn <= 225 -> span <= 256
  const int lanes = 256 / span, sub = threadIdx.x / span;
  const int e = threadIdx.x - sub * span;
  if (e >= nn) return;
  const int i = e / n, j = e - i * n;
  const cx<T> ct = conj(O[(size_t)j * n + i]);
  const long long b0 = (long long)blockIdx.x * kBcastMats;
  const long long b1 = b0 + kBcastMats < batch ? b0 + kBcastMats : batch;
#pragma unroll 4
  for (long long b = b0 + sub; b < b1; b += lanes) rhobar[(size_t)b * nn + e] = gbar[b] * ct;
}

// f[b] = |a^H b|^2 ; bwd: abar = 2 g conj(z) b, bbar = 2 g z a, z = a^H b
template <typename T, bool BWD>
__global__ void fidelity_ket_kernel(const cx<T>* __restrict__ A, const cx<T>* __restrict__ B, T* __restrict__ out,
                                    const T* __restrict__ gbar, cx<T>* __restrict__ abar, cx<T>* __restrict__ bbar,
                                    long long batch, int n, int tps) {
  const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long b = gt / tps;
  const int l = (int)(gt % tps);
  const bool live = b < batch;
  cx<T> z(T(0), T(0));
  const cx<T>* a = A + (size_t)(live ? b : 0) * n;
  const cx<T>* bb = B + (size_t)(live ? b : 0) * n;
  if (live) for (int i = l; i < n; i += tps) cfma_conj(z, a[i], bb[i]);
  z.re = group_sum(z.re, tps); z.im = group_sum(z.im, tps);
  if (!live) return;
  if (!BWD) { if (l == 0) out[b] = z.re * z.re + z.im * z.im; return; }
  const T g2 = T(2) * gbar[b];
  for (int i = l; i < n; i += tps) {
    if (abar) abar[(size_t)b * n + i] = g2 * (conj(z) * bb[i]);
    if (bbar) bbar[(size_t)b * n + i] = g2 * (z * a[i]);
  }
}

// n >= 16 with 16-byte units (one complex128, two complex64): a lane loads two units of each ket before it uses
// either, which doubles the bytes in flight per thread over the kernel above (the c64 n = 32 case had 16).
template <typename T> struct Unit;
template <> struct Unit<double> { typedef double2 V; static constexpr int E = 1; };
template <> struct Unit<float> { typedef float4 V; static constexpr int E = 2; };
__device__ __forceinline__ void unit_dot(cx<double>& z, const double2& a, const double2& b) {
  cfma_conj(z, cx<double>(a.x, a.y), cx<double>(b.x, b.y));
}
__device__ __forceinline__ void unit_dot(cx<float>& z, const float4& a, const float4& b) {
  cfma_conj(z, cx<float>(a.x, a.y), cx<float>(b.x, b.y));
  cfma_conj(z, cx<float>(a.z, a.w), cx<float>(b.z, b.w));
}
__device__ __forceinline__ double2 unit_scale(const cx<double>& w, const double2& a) {
  const cx<double> r = w * cx<double>(a.x, a.y);
  return make_double2(r.re, r.im);
}
__device__ __forceinline__ float4 unit_scale(const cx<float>& w, const float4& a) {
  const cx<float> r0 = w * cx<float>(a.x, a.y), r1 = w * cx<float>(a.z, a.w);
  return make_float4(r0.re, r0.im, r1.re, r1.im);
}
template <typename T, bool BWD>
__global__ void fidelity_ket_wide_kernel(const cx<T>* __restrict__ A, const cx<T>* __restrict__ B, T* __restrict__ out,
                                         const T* __restrict__ gbar, cx<T>* __restrict__ abar, cx<T>* __restrict__ bbar,
                                         long long batch, int units, int tps) {
  typedef typename Unit<T>::V V;
  const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long b = gt / tps;
  const int l = (int)(gt % tps);
  const bool live = b < batch;
  const V* a = reinterpret_cast<const V*>(A) + (size_t)(live ? b : 0) * units;
  const V* c = reinterpret_cast<const V*>(B) + (size_t)(live ? b : 0) * units;
  cx<T> z(T(0), T(0));
  if (live)
    for (int i = l; i < units; i += 2 * tps) {
      const bool two = i + tps < units;
      const V a0 = a[i], c0 = c[i];
      const V a1 = two ? a[i + tps] : V(), c1 = two ? c[i + tps] : V();
      unit_dot(z, a0, c0);
      if (two) unit_dot(z, a1, c1);
    }
  z.re = group_sum(z.re, tps); z.im = group_sum(z.im, tps);
  if (!live) return;
  if (!BWD) { if (l == 0) out[b] = z.re * z.re + z.im * z.im; return; }
  const T g2 = T(2) * gbar[b];
  const cx<T> wa = g2 * conj(z), wb = g2 * z;
  V* ab = reinterpret_cast<V*>(abar) + (size_t)b * units;
  V* bb = reinterpret_cast<V*>(bbar) + (size_t)b * units;
  for (int i = l; i < units; i += tps) {
    if (abar) ab[i] = unit_scale(wa, c[i]);
    if (bbar) bb[i] = unit_scale(wb, a[i]);
  }
}

// Reverse mode for long kets (256 < units <= 256 * U): one CTA per sample holds both kets in registers between
// the reduction and the cotangent stores.  The sub-warp kernel above re-reads them, and with 64 warps per SM each
// in the middle of its own 2 x 16-32 KB sample the re-read misses L1 and L2: 6 passes over HBM instead of 4.
template <typename T, int U>
__global__ void __launch_bounds__(256)
fidelity_ket_cta_bwd_kernel(const cx<T>* __restrict__ A, const cx<T>* __restrict__ B, const T* __restrict__ gbar,
                            cx<T>* __restrict__ abar, cx<T>* __restrict__ bbar, int units) {
  typedef typename Unit<T>::V V;
  __shared__ T red[2][8];
  const size_t b = blockIdx.x;
  const V* a = reinterpret_cast<const V*>(A) + b * units;
  const V* c = reinterpret_cast<const V*>(B) + b * units;
  V av[U], cv[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int i = threadIdx.x + 256 * u;
    if (i < units) { av[u] = a[i]; cv[u] = c[i]; }
  }
  cx<T> z(T(0), T(0));
#pragma unroll
  for (int u = 0; u < U; ++u)
    if ((int)threadIdx.x + 256 * u < units) unit_dot(z, av[u], cv[u]);
  z.re = warp_sum(z.re); z.im = warp_sum(z.im);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = z.re; red[1][threadIdx.x >> 5] = z.im; }
  __syncthreads();
  z = cx<T>(T(0), T(0));
#pragma unroll
  for (int w = 0; w < 8; ++w) { z.re += red[0][w]; z.im += red[1][w]; }
  const T g2 = T(2) * gbar[b];
  const cx<T> wa = g2 * conj(z), wb = g2 * z;
  V* ab = reinterpret_cast<V*>(abar) + b * units;
  V* bb = reinterpret_cast<V*>(bbar) + b * units;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int i = threadIdx.x + 256 * u;
    if (i < units) {
      if (abar) ab[i] = unit_scale(wa, cv[u]);
      if (bbar) bb[i] = unit_scale(wb, av[u]);
    }
  }
}

// F[b] = sqrt(1 - t^2), t = 0.5 sum_i |rho1_ii - rho2_ii|   (elementwise-abs surrogate, as written)
template <typename T, bool BWD>
__global__ void fidelity_dm_kernel(const cx<T>* __restrict__ R1, const cx<T>* __restrict__ R2, T* __restrict__ out,
                                   const T* __restrict__ gbar, cx<T>* __restrict__ r1bar, cx<T>* __restrict__ r2bar,
                                   long long batch, int n, int tps) {
  const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long b = gt / tps;
  const int l = (int)(gt % tps);
  const bool live = b < batch;
  const cx<T>* r1 = R1 + (size_t)(live ? b : 0) * n * n;
  const cx<T>* r2 = R2 + (size_t)(live ? b : 0) * n * n;
  T t = T(0);
  if (live) for (int i = l; i < n; i += tps) t += cabs_(r1[(size_t)i * n + i] - r2[(size_t)i * n + i]);
  t = T(0.5) * group_sum(t, tps);
  if (!live) return;
  const T F = sqrt(T(1) - t * t);
  if (!BWD) { if (l == 0) out[b] = F; return; }
  // dF/dt = -t/F ; d|d_i|/d(d_i) in the conjugate convention = d_i/|d_i| (0 at d_i = 0)
  const T coef = gbar[b] * (-t / F) * T(0.5);
  for (int i = l; i < n; i += tps) {
    const cx<T> d = r1[(size_t)i * n + i] - r2[(size_t)i * n + i];
    const T ad = cabs_(d);
    const cx<T> v = ad > T(0) ? (coef / ad) * d : cx<T>(T(0), T(0));
    for (int j = 0; j < n; ++j) {
      const cx<T> w = (j == i) ? v : cx<T>(T(0), T(0));
      if (r1bar) r1bar[(size_t)b * n * n + (size_t)i * n + j] = w;
      if (r2bar) r2bar[(size_t)b * n * n + (size_t)i * n + j] = -w;
    }
  }
}

// ------------------------------------------------------------------ n <= 8: one THREAD per sample
// A ket of n <= 8 amplitudes is 16-128 contiguous bytes: one thread loads it with 16-byte vector loads
// (every fetched sector is fully used), keeps it in registers, and writes its results the same way.
// The configs that live here: Qubit_Rotation (n = 2, 10^6 parameter sets), Circuit-Learning (n = 4),
// Unitary-Learning (n = 8).  A shared operator sits in shared memory (broadcast reads).
template <typename T, int N>
__device__ __forceinline__ void load_vec(cx<T> (&v)[N], const cx<T>* __restrict__ src) {
  if (sizeof(T) == 4 && N % 2 == 0) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
      const float4 q = reinterpret_cast<const float4*>(src)[i];
      v[2 * i] = cx<T>((T)q.x, (T)q.y); v[2 * i + 1] = cx<T>((T)q.z, (T)q.w);
    }
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = src[i];
  }
}
template <typename T, int N>
__device__ __forceinline__ void store_vec(cx<T>* __restrict__ dst, const cx<T> (&v)[N]) {
  if (sizeof(T) == 4 && N % 2 == 0) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i)
      reinterpret_cast<float4*>(dst)[i] = make_float4((float)v[2 * i].re, (float)v[2 * i].im, (float)v[2 * i + 1].re, (float)v[2 * i + 1].im);
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) dst[i] = v[i];
  }
}

// shared operator (ob = 0): forward value, or (BWD) ketbar = conj(g) O psi + g O^H psi
template <typename T, int N, bool BWD>
__global__ void expect_ket_small_kernel(const cx<T>* __restrict__ O, const cx<T>* __restrict__ psi, cx<T>* __restrict__ out,
                                        const cx<T>* __restrict__ gbar, cx<T>* __restrict__ ketbar, long long batch) {
  __shared__ cx<T> sO[N * N];
  for (int e = threadIdx.x; e < N * N; e += blockDim.x) sO[e] = O[e];
  __syncthreads();
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  cx<T> p[N];
  load_vec<T, N>(p, psi + (size_t)b * N);
  if (!BWD) {
    cx<T> acc(T(0), T(0));
#pragma unroll
    for (int i = 0; i < N; ++i) {
      cx<T> row(T(0), T(0));
#pragma unroll
      for (int j = 0; j < N; ++j) cfma(row, sO[i * N + j], p[j]);
      cfma_conj(acc, p[i], row);
    }
    out[b] = acc;
  } else {
    const cx<T> g = gbar[b];
    cx<T> kb[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      cx<T> r1(T(0), T(0)), r2(T(0), T(0));
#pragma unroll
      for (int j = 0; j < N; ++j) { cfma(r1, sO[i * N + j], p[j]); cfma_conj(r2, sO[j * N + i], p[j]); }
      kb[i] = conj(g) * r1 + g * r2;
    }
    store_vec<T, N>(ketbar + (size_t)b * N, kb);
  }
}

template <typename T, int N, bool BWD>
__global__ void fidelity_ket_small_kernel(const cx<T>* __restrict__ A, const cx<T>* __restrict__ B, T* __restrict__ out,
                                          const T* __restrict__ gbar, cx<T>* __restrict__ abar, cx<T>* __restrict__ bbar,
                                          long long batch) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  cx<T> a[N], c[N];
  load_vec<T, N>(a, A + (size_t)b * N);
  load_vec<T, N>(c, B + (size_t)b * N);
  cx<T> z(T(0), T(0));
#pragma unroll
  for (int i = 0; i < N; ++i) cfma_conj(z, a[i], c[i]);
  if (!BWD) { out[b] = z.re * z.re + z.im * z.im; return; }
  const T g2 = T(2) * gbar[b];
  cx<T> o[N];
  if (abar) {
#pragma unroll
    for (int i = 0; i < N; ++i) o[i] = g2 * (conj(z) * c[i]);
    store_vec<T, N>(abar + (size_t)b * N, o);
  }
  if (bbar) {
#pragma unroll
    for (int i = 0; i < N; ++i) o[i] = g2 * (z * a[i]);
    store_vec<T, N>(bbar + (size_t)b * N, o);
  }
}

template <typename T, int N>
static int run_small(int op, const void* p0, const void* p1, const void* p2, void* o0, void* o1, int64_t batch, cudaStream_t st) {
  const int threads = 256;
  const unsigned grid = (unsigned)((batch + threads - 1) / threads);
  switch (op) {
    case 0: expect_ket_small_kernel<T, N, false><<<grid, threads, 0, st>>>((const cx<T>*)p0, (const cx<T>*)p1, (cx<T>*)o0, nullptr, nullptr, batch); break;
    case 1: expect_ket_small_kernel<T, N, true><<<grid, threads, 0, st>>>((const cx<T>*)p0, (const cx<T>*)p1, nullptr, (const cx<T>*)p2, (cx<T>*)o0, batch); break;
    case 4: fidelity_ket_small_kernel<T, N, false><<<grid, threads, 0, st>>>((const cx<T>*)p0, (const cx<T>*)p1, (T*)o0, nullptr, nullptr, nullptr, batch); break;
    case 5: fidelity_ket_small_kernel<T, N, true><<<grid, threads, 0, st>>>((const cx<T>*)p0, (const cx<T>*)p1, nullptr, (const T*)p2, (cx<T>*)o0, (cx<T>*)o1, batch); break;
    default: return QG_ERR_ARG;
  }
  QG_CHECK_LAUNCH();
  return QG_OK;
}

static inline unsigned grid_for(int64_t batch, int tps, int threads) {
  const int64_t total = batch * tps;
  return (unsigned)((total + threads - 1) / threads);
}

template <typename T>
static int run(int op, const void* p0, const void* p1, const void* p2, void* o0, void* o1, int64_t batch, int n, int ob,
               cudaStream_t st) {
  // thread-per-sample fast paths: expect (ket branch, shared operator, ketbar only) and ket fidelity at n = 2, 4, 8
  // 16-byte vector accesses need 16-byte aligned operands (a complex64 view may start on an odd element)
  const bool al16 = (((uintptr_t)p0 | (uintptr_t)p1 | (uintptr_t)o0 | (uintptr_t)o1) & 15) == 0;
  const bool small_ok = al16 && (n == 2 || n == 4 || (n == 8 && sizeof(T) == 4)) &&
                        ((op == 0 && !ob) || (op == 1 && !ob && o0 && !o1) || op == 4 || op == 5);
  if (small_ok) {
    if (n == 2) return run_small<T, 2>(op, p0, p1, p2, o0, o1, batch, st);
    if (n == 4) return run_small<T, 4>(op, p0, p1, p2, o0, o1, batch, st);
    return run_small<T, 8>(op, p0, p1, p2, o0, o1, batch, st);
  }
  const int tps = group_size(n), threads = 256;
  const unsigned grid = grid_for(batch, tps, threads);
  switch (op) {
    case 0: { const int t2 = (ob && n >= 16) ? 32 : tps; expect_ket_fwd_kernel<T><<<grid_for(batch, t2, threads), threads, 0, st>>>((const cx<T>*)p0, (const cx<T>*)p1, (cx<T>*)o0, batch, n, ob, t2); break; }
    case 1:
      if (o1 && !ob) QG_RETURN_IF_CUDA(cudaMemsetAsync(o1, 0, (size_t)n * n * sizeof(cx<T>), st));
      expect_ket_bwd_kernel<T><<<grid, threads, 0, st>>>((const cx<T>*)p0, (const cx<T>*)p1, (const cx<T>*)p2, (cx<T>*)o0, (cx<T>*)o1, batch, n, ob, tps); break;
    case 2: {
      const int t2 = n >= 16 ? 32 : tps;
      const size_t sm = (!ob && n >= 16 && (size_t)n * n * sizeof(cx<T>) <= 48 * 1024) ? (size_t)n * n * sizeof(cx<T>) : 0;
      expect_dm_fwd_kernel<T><<<grid_for(batch, t2, threads), threads, sm, st>>>((const cx<T>*)p0, (const cx<T>*)p1, (cx<T>*)o0, batch, n, ob, t2, sm ? 1 : 0);
      break;
    }
    case 3:
      if (!ob && o0 && !o1 && n < 16) {
        expect_dm_bwd_bcast_kernel<T><<<(unsigned)((batch + kBcastMats - 1) / kBcastMats), threads, 0, st>>>((const cx<T>*)p0, (const cx<T>*)p2, (cx<T>*)o0, batch, n);
        break;
      }
      if (o1 && !ob) QG_RETURN_IF_CUDA(cudaMemsetAsync(o1, 0, (size_t)n * n * sizeof(cx<T>), st));
      {
        const int t2 = n >= 16 ? 32 : tps;
        const size_t sm = (!ob && n >= 16 && (size_t)n * n * sizeof(cx<T>) <= 48 * 1024) ? (size_t)n * n * sizeof(cx<T>) : 0;
        expect_dm_bwd_kernel<T><<<grid_for(batch, t2, threads), threads, sm, st>>>((const cx<T>*)p0, (const cx<T>*)p1, (const cx<T>*)p2, (cx<T>*)o0, (cx<T>*)o1, batch, n, ob, t2, sm ? 1 : 0);
      }
      break;
    case 4: case 5:
      if (al16 && n >= 16 && n % Unit<T>::E == 0) {
        const int units = n / Unit<T>::E;
        if (op == 5 && units > 256 && units <= 2048 && batch <= 0x7fffffffLL) {
          if (units <= 1024) fidelity_ket_cta_bwd_kernel<T, 4><<<(unsigned)batch, 256, 0, st>>>((const cx<T>*)p0, (const cx<T>*)p1, (const T*)p2, (cx<T>*)o0, (cx<T>*)o1, units);
          else fidelity_ket_cta_bwd_kernel<T, 8><<<(unsigned)batch, 256, 0, st>>>((const cx<T>*)p0, (const cx<T>*)p1, (const T*)p2, (cx<T>*)o0, (cx<T>*)o1, units);
          break;
        }
        const int t2 = group_size(units) >= 2 ? group_size(units) / 2 : 1;       // two units per lane and pass
        if (op == 4) fidelity_ket_wide_kernel<T, false><<<grid_for(batch, t2, threads), threads, 0, st>>>((const cx<T>*)p0, (const cx<T>*)p1, (T*)o0, nullptr, nullptr, nullptr, batch, units, t2);
        else fidelity_ket_wide_kernel<T, true><<<grid_for(batch, t2, threads), threads, 0, st>>>((const cx<T>*)p0, (const cx<T>*)p1, nullptr, (const T*)p2, (cx<T>*)o0, (cx<T>*)o1, batch, units, t2);
        break;
      }
      if (op == 5) { fidelity_ket_kernel<T, true><<<grid, threads, 0, st>>>((const cx<T>*)p0, (const cx<T>*)p1, nullptr, (const T*)p2, (cx<T>*)o0, (cx<T>*)o1, batch, n, tps); break; }
      fidelity_ket_kernel<T, false><<<grid, threads, 0, st>>>((const cx<T>*)p0, (const cx<T>*)p1, (T*)o0, nullptr, nullptr, nullptr, batch, n, tps); break;
    case 6: fidelity_dm_kernel<T, false><<<grid, threads, 0, st>>>((const cx<T>*)p0, (const cx<T>*)p1, (T*)o0, nullptr, nullptr, nullptr, batch, n, tps); break;
    case 7: fidelity_dm_kernel<T, true><<<grid, threads, 0, st>>>((const cx<T>*)p0, (const cx<T>*)p1, nullptr, (const T*)p2, (cx<T>*)o0, (cx<T>*)o1, batch, n, tps); break;
    default: return QG_ERR_ARG;
  }
  QG_CHECK_LAUNCH();
  return QG_OK;
}

}  // namespace state

int state_op(int dtype, int op, const void* p0, const void* p1, const void* p2, void* o0, void* o1, int64_t batch,
             int n, int ob, cudaStream_t st) {
  return dtype == QG_C128 ? state::run<double>(op, p0, p1, p2, o0, o1, batch, n, ob, st)
                          : state::run<float>(op, p0, p1, p2, o0, o1, batch, n, ob, st);
}

}  // namespace qg

// K5 / K6 -- operator builders on the device, forward and reverse mode.
//   Displace.__call__ / coherent      qgrad/qgrad_qutip.py:216-261, 302-316
//   _make_rot, Unitary.__call__       qgrad/qgrad_qutip.py:424-459, 462-555
//   rot (ZYZ), Qubit_Rotation cost    examples/Qubit_Rotation.py:40-92
// These are O(n^2)-O(n^3) flops on O(n^2) bytes for n <= ~50, or O(1) per parameter set: the
// batched forms are HBM / launch bound, so each sample is owned by one CTA or one warp and
// nothing round-trips through HBM between the build and its gradient.
#include "common.cuh"

namespace qg {
namespace build {

template <typename T> __device__ __forceinline__ cx<T> expi(T x) { T s, c; sincos_(x, &s, &c); return cx<T>(c, s); }
// s_j = i^(j % 2)
template <typename T> __device__ __forceinline__ cx<T> tscale(int j) { return (j & 1) ? cx<T>(T(0), T(1)) : cx<T>(T(1), T(0)); }

template <typename T> __device__ __forceinline__ T block_sum(T v, T* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  T tot = T(0);
  for (int w = 0; w < (int)(blockDim.x + 31) / 32; ++w) tot += red[w];
  return tot;
}

// ------------------------------------------------------------------ Displace (matrix form)
// smem: V (n*n reals) | w (n complex) | iw = i*lambda*w (n complex) | p (n complex) | red (32)
template <typename T, bool BWD>
__global__ void displace_kernel(const T* __restrict__ V, const T* __restrict__ lam, const cx<T>* __restrict__ alpha,
                                cx<T>* __restrict__ D, const cx<T>* __restrict__ Dbar, cx<T>* __restrict__ alphabar,
                                int n) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sV = reinterpret_cast<T*>(smem_raw);
  cx<T>* sw = reinterpret_cast<cx<T>*>(sV + (size_t)n * n + ((n * n) & 1));
  cx<T>* siw = sw + n;
  cx<T>* sp = siw + n;
  T* red = reinterpret_cast<T*>(sp + n);
  const long long b = blockIdx.x;
  const cx<T> a = alpha[b];
  const T r = cabs_(a);
  const T phi = (r == T(0)) ? T(0) : atan2(a.im, a.re);
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) sV[e] = V[e];
  for (int m = threadIdx.x; m < n; m += blockDim.x) {
    const cx<T> w = expi<T>(r * lam[m]);
    sw[m] = w;
    siw[m] = cx<T>(-lam[m] * w.im, lam[m] * w.re);                 // i*lambda*w
    sp[m] = tscale<T>(m) * expi<T>(-phi * (T)m);                   // s_m e^{-i phi m}
  }
  __syncthreads();
  T rbar = T(0), phibar = T(0);
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    const int j = e / n, k = e % n;
    cx<T> M(T(0), T(0)), N(T(0), T(0));
    for (int m = 0; m < n; ++m) {
      const T vv = sV[j * n + m] * sV[k * n + m];
      M.re = fma(vv, sw[m].re, M.re); M.im = fma(vv, sw[m].im, M.im);
      if (BWD) { N.re = fma(vv, siw[m].re, N.re); N.im = fma(vv, siw[m].im, N.im); }
    }
    const cx<T> ph = conj(sp[j]) * sp[k];
    const cx<T> d = ph * M;
    if (!BWD) {
      D[(size_t)b * n * n + e] = d;
    } else {
      const cx<T> g = Dbar[(size_t)b * n * n + e];
      const cx<T> dr = ph * N;                                     // dD/dr
      rbar += g.re * dr.re + g.im * dr.im;                         // Re(conj(g) dr)
      // dD/dphi = i (j-k) D :  Re(conj(g) * i (j-k) d) = -(j-k) Im(conj(g) d)
      phibar += -(T)(j - k) * (g.re * d.im - g.im * d.re);
    }
  }
  if (BWD) {
    rbar = block_sum(rbar, red);
    phibar = block_sum(phibar, red);
    if (threadIdx.x == 0) {
      cx<T> ab(T(0), T(0));
      if (r > T(0)) ab = cx<T>(rbar * a.re / r - phibar * a.im / (r * r), rbar * a.im / r + phibar * a.re / (r * r));
      alphabar[b] = ab;
    }
  }
}

// ------------------------------------------------------------------ Displace (matrix form), register tiled
// M = V diag(w) V^T is an n^3 product per alpha: above n ~ 20 the build is FMA-pipe bound, not HBM bound.
// Each thread owns a 4 x 4 block of M (and, BWD, of N = V diag(i lambda w) V^T) in registers; per m it
// reads two 4-vectors of V^T (16-byte loads; V^T is staged once per CTA, zero padded to a multiple of 4,
// and shared by all alphas of the CTA) and w_m: 32 FMAs per 3 shared-memory loads.  (np/4)^2 threads per
// alpha, 128 / that alphas per CTA.
// smem: Vt [np][np] reals | per alpha: w [np], iw [np], p [np] complex | red [128][2] reals
template <typename T, bool BWD>
__global__ void __launch_bounds__(128) displace_tiled_kernel(const T* __restrict__ V, const T* __restrict__ lam,
                                                             const cx<T>* __restrict__ alpha, cx<T>* __restrict__ D,
                                                             const cx<T>* __restrict__ Dbar, cx<T>* __restrict__ alphabar,
                                                             long long batch, int n, int np, int apb) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sVt = reinterpret_cast<T*>(smem_raw);
  cx<T>* vecs = reinterpret_cast<cx<T>*>(sVt + (size_t)np * np);
  T* red = reinterpret_cast<T*>(vecs + (size_t)apb * 3 * np);
  const int nb = np / 4, tpa = nb * nb;
  const int tid = threadIdx.x;
  for (int e = tid; e < np * np; e += blockDim.x) {
    const int m = e / np, j = e - m * np;
    sVt[e] = (m < n && j < n) ? V[(size_t)j * n + m] : T(0);
  }
  const int a_loc = tid / tpa, blk = tid - a_loc * tpa;
  const bool worker = a_loc < apb;
  const long long b = (long long)blockIdx.x * apb + a_loc;
  const bool live = worker && b < batch;
  cx<T>* sw = vecs + (size_t)(worker ? a_loc : 0) * 3 * np;
  cx<T>* siw = sw + np;
  cx<T>* sp = siw + np;
  cx<T> a(T(0), T(0));
  T r = T(0), phi = T(0);
  if (live) {
    a = alpha[b];
    r = cabs_(a);
    phi = (r == T(0)) ? T(0) : atan2(a.im, a.re);
    for (int m = blk; m < np; m += tpa) {
      const T lm = m < n ? lam[m] : T(0);
      const cx<T> w = expi<T>(r * lm);
      sw[m] = w;
      siw[m] = cx<T>(-lm * w.im, lm * w.re);                        // i*lambda*w
      sp[m] = tscale<T>(m) * expi<T>(-phi * (T)m);                  // s_m e^{-i phi m}
    }
  }
  __syncthreads();
  T rbar = T(0), phibar = T(0);
  if (live) {
    const int bj = blk / nb, bk = blk - bj * nb;
    cx<T> M[4][4], N[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) { M[u][v] = cx<T>(T(0), T(0)); N[u][v] = cx<T>(T(0), T(0)); }
    for (int m = 0; m < n; ++m) {
      T vj[4], vk[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { vj[u] = sVt[m * np + bj * 4 + u]; vk[u] = sVt[m * np + bk * 4 + u]; }
      const cx<T> w = sw[m];
      cx<T> wk[4], iwk[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) wk[v] = cx<T>(vk[v] * w.re, vk[v] * w.im);
      if (BWD) {
        const cx<T> iw = siw[m];
#pragma unroll
        for (int v = 0; v < 4; ++v) iwk[v] = cx<T>(vk[v] * iw.re, vk[v] * iw.im);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          M[u][v].re = fma(vj[u], wk[v].re, M[u][v].re); M[u][v].im = fma(vj[u], wk[v].im, M[u][v].im);
          if (BWD) { N[u][v].re = fma(vj[u], iwk[v].re, N[u][v].re); N[u][v].im = fma(vj[u], iwk[v].im, N[u][v].im); }
        }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = bj * 4 + u;
      if (j >= n) continue;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int k = bk * 4 + v;
        if (k >= n) continue;
        const cx<T> ph = conj(sp[j]) * sp[k];
        const cx<T> d = ph * M[u][v];
        const size_t e = (size_t)b * n * n + (size_t)j * n + k;
        if (!BWD) {
          D[e] = d;
        } else {
          const cx<T> g = Dbar[e];
          const cx<T> dr = ph * N[u][v];                            // dD/dr
          rbar += g.re * dr.re + g.im * dr.im;                      // Re(conj(g) dr)
          phibar += -(T)(j - k) * (g.re * d.im - g.im * d.re);      // dD/dphi = i (j-k) D
        }
      }
    }
  }
  if (BWD) {
    // per-alpha reduction over its tpa threads, in thread order (deterministic)
    red[2 * tid] = rbar; red[2 * tid + 1] = phibar;
    __syncthreads();
    if (live && blk == 0) {
      T rb = T(0), pb = T(0);
      for (int q = 0; q < tpa; ++q) { rb += red[2 * (tid + q)]; pb += red[2 * (tid + q) + 1]; }
      cx<T> ab(T(0), T(0));
      if (r > T(0)) ab = cx<T>(rb * a.re / r - pb * a.im / (r * r), rb * a.im / r + pb * a.re / (r * r));
      alphabar[b] = ab;
    }
  }
}

// ------------------------------------------------------------------ Displace applied to a ket
// y = conj(p) o V (w o V^T (p o x)).  One CTA per sample; vectors in smem.
// BWD: xbar = conj(p) o V (conj(w) o V^T (p o ybar)) and alphabar from
//   dy/dr = conj(p) o V (i lambda w o u),  dy/dphi = i (j o y) - i D (k o x).
template <typename T, bool BWD>
__global__ void displace_apply_kernel(const T* __restrict__ V, const T* __restrict__ lam, const cx<T>* __restrict__ alpha,
                                      const cx<T>* __restrict__ X, cx<T>* __restrict__ Y, const cx<T>* __restrict__ Ybar,
                                      cx<T>* __restrict__ alphabar, cx<T>* __restrict__ Xbar, int n) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<T>* sp = reinterpret_cast<cx<T>*>(smem_raw);   // p
  cx<T>* sw = sp + n;                               // w
  cx<T>* t0 = sw + n;                               // scratch vectors
  cx<T>* t1 = t0 + n;
  cx<T>* t2 = t1 + n;
  cx<T>* t3 = t2 + n;
  T* red = reinterpret_cast<T*>(t3 + n);
  const long long b = blockIdx.x;
  const cx<T> a = alpha[b];
  const T r = cabs_(a);
  const T phi = (r == T(0)) ? T(0) : atan2(a.im, a.re);
  const cx<T>* x = X + (size_t)b * n;
  for (int m = threadIdx.x; m < n; m += blockDim.x) {
    sw[m] = expi<T>(r * lam[m]);
    sp[m] = tscale<T>(m) * expi<T>(-phi * (T)m);
    t0[m] = sp[m] * x[m];                                           // p o x
    if (BWD) t2[m] = (T)m * t0[m];                                  // p o (k x)
  }
  __syncthreads();
  // u = V^T t0 (and u' = V^T t2), scaled by w
  for (int m = threadIdx.x; m < n; m += blockDim.x) {
    cx<T> u(T(0), T(0)), u2(T(0), T(0));
    for (int k = 0; k < n; ++k) {
      const T v = V[(size_t)k * n + m];
      u.re = fma(v, t0[k].re, u.re); u.im = fma(v, t0[k].im, u.im);
      if (BWD) { u2.re = fma(v, t2[k].re, u2.re); u2.im = fma(v, t2[k].im, u2.im); }
    }
    t1[m] = sw[m] * u;
    if (BWD) t3[m] = sw[m] * u2;
  }
  __syncthreads();
  T rbar = T(0), phibar = T(0);
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    cx<T> y(T(0), T(0)), yr(T(0), T(0)), y2(T(0), T(0));
    for (int m = 0; m < n; ++m) {
      const T v = V[(size_t)j * n + m];
      y.re = fma(v, t1[m].re, y.re); y.im = fma(v, t1[m].im, y.im);
      if (BWD) {
        const T vl = v * lam[m];
        yr.re = fma(-vl, t1[m].im, yr.re); yr.im = fma(vl, t1[m].re, yr.im);   // i lambda w u
        y2.re = fma(v, t3[m].re, y2.re); y2.im = fma(v, t3[m].im, y2.im);
      }
    }
    const cx<T> cp = conj(sp[j]);
    y = cp * y;
    if (!BWD) {
      Y[(size_t)b * n + j] = y;
    } else {
      yr = cp * yr; y2 = cp * y2;
      const cx<T> g = Ybar[(size_t)b * n + j];
      rbar += g.re * yr.re + g.im * yr.im;
      // dy/dphi = i (j y - y2):  Re(conj(g) i z) = -Im(conj(g) z), z = j y - y2
      const cx<T> z((T)j * y.re - y2.re, (T)j * y.im - y2.im);
      phibar += -(g.re * z.im - g.im * z.re);
    }
  }
  if (BWD) {
    rbar = block_sum(rbar, red);
    phibar = block_sum(phibar, red);
    if (threadIdx.x == 0 && alphabar) {
      cx<T> ab(T(0), T(0));
      if (r > T(0)) ab = cx<T>(rbar * a.re / r - phibar * a.im / (r * r), rbar * a.im / r + phibar * a.re / (r * r));
      alphabar[b] = ab;
    }
    if (Xbar) {
      // xbar = D^H ybar = conj(p) o V (conj(w) o V^T (p o ybar))
      __syncthreads();
      for (int m = threadIdx.x; m < n; m += blockDim.x) t0[m] = sp[m] * Ybar[(size_t)b * n + m];
      __syncthreads();
      for (int m = threadIdx.x; m < n; m += blockDim.x) {
        cx<T> u(T(0), T(0));
        for (int k = 0; k < n; ++k) { const T v = V[(size_t)k * n + m]; u.re = fma(v, t0[k].re, u.re); u.im = fma(v, t0[k].im, u.im); }
        t1[m] = conj(sw[m]) * u;
      }
      __syncthreads();
      for (int j = threadIdx.x; j < n; j += blockDim.x) {
        cx<T> y(T(0), T(0));
        for (int m = 0; m < n; ++m) { const T v = V[(size_t)j * n + m]; y.re = fma(v, t1[m].re, y.re); y.im = fma(v, t1[m].im, y.im); }
        Xbar[(size_t)b * n + j] = conj(sp[j]) * y;
      }
    }
  }
}

// ------------------------------------------------------------------ Displace applied to kets, 32 samples per CTA
// lane = sample.  V (row major, for V^T t) and V^T (for V u) sit in shared memory once per CTA, padded to
// a row length np that is a multiple of 4; vectors are stored [index][sample] with a pitch of 33 so that
// the lane-indexed accesses are conflict free.  A warp owns blocks of four output indices: per inner step
// one lane-indexed vector element and one broadcast 4-vector of V feed eight FMAs.  Global I/O is
// staged through the same buffers so that it is coalesced (a CTA's 32 kets are contiguous in memory).
//   y = conj(p) o V (w o V^T (p o x)),  p_m = s_m e^{-i phi m},  w_m = e^{i r lambda_m},  alpha = r e^{i phi}
//   BWD: xbar = conj(p) o V (conj(w) o V^T (p o ybar));  rbar = Re <ybar, dy/dr>, phibar = Re <ybar, dy/dphi>,
//        dy/dr = conj(p) o V (i lambda w o u'),  dy/dphi = i (j o y) - i conj(p) o V (w o V^T (k o p o x)).
constexpr int kApplyPitch = 33;
template <typename T, bool BWD>
__global__ void __launch_bounds__(128) displace_apply_batched_kernel(
    const T* __restrict__ V, const T* __restrict__ lam, const cx<T>* __restrict__ alpha, const cx<T>* __restrict__ X,
    cx<T>* __restrict__ Y, const cx<T>* __restrict__ Ybar, cx<T>* __restrict__ alphabar, cx<T>* __restrict__ Xbar,
    long long batch, int n, int np) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sV = reinterpret_cast<T*>(smem_raw);                 // [n][np]   sV[k][m]  = V[k][m]
  T* sVt = sV + (size_t)n * np;                           // [n][np]   sVt[m][j] = V[j][m]
  T* slam = sVt + (size_t)n * np;                         // [np]
  T* srp = slam + np;                                     // r[32], phi[32]
  T* sred = srp + 64;                                     // [4][32][2]
  cx<T>* buf = reinterpret_cast<cx<T>*>(sred + 256);
  const size_t VB = (size_t)np * kApplyPitch;             // one vector buffer
  cx<T>* bP = buf;                                        // p
  cx<T>* bT0 = buf + VB;                                  // p o x            (later: p o ybar)
  cx<T>* bU = buf + 2 * VB;                               // w o V^T t0       (later: conj(w) o V^T (p o ybar))
  cx<T>* bT2 = buf + 3 * VB;                              // BWD: k o t0      (later: y, for the coalesced store: fwd uses bT0)
  cx<T>* bU2 = buf + 4 * VB;                              // BWD: w o V^T t2
  cx<T>* bW = buf + 5 * VB;                               // BWD: w
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long b0 = (long long)blockIdx.x * 32;
  const int ns = (int)((batch - b0 < 32) ? batch - b0 : 32);
  for (int e = tid; e < n * np; e += 128) {
    const int k = e / np, m = e - k * np;
    sV[e] = m < n ? V[(size_t)k * n + m] : T(0);
    sVt[e] = m < n ? V[(size_t)m * n + k] : T(0);
  }
  for (int m = tid; m < np; m += 128) slam[m] = m < n ? lam[m] : T(0);
  if (tid < 32) {
    T r = T(0), phi = T(0);
    if (tid < ns) { const cx<T> a = alpha[b0 + tid]; r = cabs_(a); phi = (r == T(0)) ? T(0) : atan2(a.im, a.re); }
    srp[tid] = r; srp[32 + tid] = phi;
  }
  __syncthreads();
  // ---- load: p, t0 = p o x (t2 = k t0); coalesced over the CTA's contiguous chunk of kets
  for (int e = tid; e < 32 * n; e += 128) {
    const int sidx_ = e / n, k = e - sidx_ * n;
    cx<T> pv(T(0), T(0)), xv(T(0), T(0));
    if (sidx_ < ns) {
      pv = tscale<T>(k) * expi<T>(-srp[32 + sidx_] * (T)k);
      xv = X[(size_t)b0 * n + e];
    }
    const cx<T> t0 = pv * xv;
    bP[k * kApplyPitch + sidx_] = pv;
    bT0[k * kApplyPitch + sidx_] = t0;
    if (BWD) bT2[k * kApplyPitch + sidx_] = cx<T>((T)k * t0.re, (T)k * t0.im);
  }
  __syncthreads();
  const T rs = srp[lane];
  const int nblk = np / 4;
  // ---- u = w o V^T t0   (u2 = w o V^T t2)
  for (int mb = warp; mb < nblk; mb += 4) {
    cx<T> acc[4], acc2[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { acc[u] = cx<T>(T(0), T(0)); acc2[u] = cx<T>(T(0), T(0)); }
    for (int k = 0; k < n; ++k) {
      const cx<T> xk = bT0[k * kApplyPitch + lane];
      T v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = sV[k * np + mb * 4 + u];
#pragma unroll
      for (int u = 0; u < 4; ++u) { acc[u].re = fma(v[u], xk.re, acc[u].re); acc[u].im = fma(v[u], xk.im, acc[u].im); }
      if (BWD) {
        const cx<T> x2 = bT2[k * kApplyPitch + lane];
#pragma unroll
        for (int u = 0; u < 4; ++u) { acc2[u].re = fma(v[u], x2.re, acc2[u].re); acc2[u].im = fma(v[u], x2.im, acc2[u].im); }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int m = mb * 4 + u;
      const cx<T> w = expi<T>(rs * slam[m]);
      bU[m * kApplyPitch + lane] = w * acc[u];
      if (BWD) { bU2[m * kApplyPitch + lane] = w * acc2[u]; bW[m * kApplyPitch + lane] = w; }
    }
  }
  __syncthreads();
  // ---- y = conj(p) o V u   (+ yr, y2 and the (r, phi) cotangents)
  cx<T>* bYout = BWD ? bT2 : bT0;   // fwd: t0 is dead after the first product; bwd: y is not stored at all
  T rbar = T(0), phibar = T(0);
  for (int jb = warp; jb < nblk; jb += 4) {
    cx<T> acc[4], accr[4], acc2[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { acc[u] = cx<T>(T(0), T(0)); accr[u] = cx<T>(T(0), T(0)); acc2[u] = cx<T>(T(0), T(0)); }
    for (int m = 0; m < n; ++m) {
      const cx<T> um = bU[m * kApplyPitch + lane];
      T v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = sVt[m * np + jb * 4 + u];
#pragma unroll
      for (int u = 0; u < 4; ++u) { acc[u].re = fma(v[u], um.re, acc[u].re); acc[u].im = fma(v[u], um.im, acc[u].im); }
      if (BWD) {
        const T lm = slam[m];
        const cx<T> iu(-lm * um.im, lm * um.re);                   // i lambda u
        const cx<T> u2 = bU2[m * kApplyPitch + lane];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          accr[u].re = fma(v[u], iu.re, accr[u].re); accr[u].im = fma(v[u], iu.im, accr[u].im);
          acc2[u].re = fma(v[u], u2.re, acc2[u].re); acc2[u].im = fma(v[u], u2.im, acc2[u].im);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = jb * 4 + u;
      if (j >= n) continue;
      const cx<T> cp = conj(bP[j * kApplyPitch + lane]);
      const cx<T> y = cp * acc[u];
      if (!BWD) {
        bYout[j * kApplyPitch + lane] = y;
      } else if (lane < ns) {
        const cx<T> yr = cp * accr[u], y2 = cp * acc2[u];
        const cx<T> g = Ybar[(size_t)(b0 + lane) * n + j];
        rbar += g.re * yr.re + g.im * yr.im;
        const cx<T> z((T)j * y.re - y2.re, (T)j * y.im - y2.im);   // dy/dphi = i z
        phibar += -(g.re * z.im - g.im * z.re);
      }
    }
  }
  if (!BWD) {
    __syncthreads();
    for (int e = tid; e < ns * n; e += 128) { const int sidx_ = e / n, k = e - sidx_ * n; Y[(size_t)b0 * n + e] = bYout[k * kApplyPitch + sidx_]; }
    return;
  }
  sred[(warp * 32 + lane) * 2] = rbar; sred[(warp * 32 + lane) * 2 + 1] = phibar;
  __syncthreads();
  if (warp == 0 && lane < ns && alphabar) {
    T rb = T(0), pb = T(0);
    for (int w = 0; w < 4; ++w) { rb += sred[(w * 32 + lane) * 2]; pb += sred[(w * 32 + lane) * 2 + 1]; }
    const cx<T> a = alpha[b0 + lane];
    const T r = rs;
    cx<T> ab(T(0), T(0));
    if (r > T(0)) ab = cx<T>(rb * a.re / r - pb * a.im / (r * r), rb * a.im / r + pb * a.re / (r * r));
    alphabar[b0 + lane] = ab;
  }
  if (!Xbar) return;
  // ---- xbar = conj(p) o V (conj(w) o V^T (p o ybar))
  for (int e = tid; e < 32 * n; e += 128) {
    const int sidx_ = e / n, k = e - sidx_ * n;
    cx<T> g(T(0), T(0));
    if (sidx_ < ns) g = Ybar[(size_t)b0 * n + e];
    bT0[k * kApplyPitch + sidx_] = bP[k * kApplyPitch + sidx_] * g;
  }
  __syncthreads();
  for (int mb = warp; mb < nblk; mb += 4) {
    cx<T> acc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) acc[u] = cx<T>(T(0), T(0));
    for (int k = 0; k < n; ++k) {
      const cx<T> xk = bT0[k * kApplyPitch + lane];
#pragma unroll
      for (int u = 0; u < 4; ++u) { const T v = sV[k * np + mb * 4 + u]; acc[u].re = fma(v, xk.re, acc[u].re); acc[u].im = fma(v, xk.im, acc[u].im); }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int m = mb * 4 + u; bU[m * kApplyPitch + lane] = conj(bW[m * kApplyPitch + lane]) * acc[u]; }
  }
  __syncthreads();
  for (int jb = warp; jb < nblk; jb += 4) {
    cx<T> acc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) acc[u] = cx<T>(T(0), T(0));
    for (int m = 0; m < n; ++m) {
      const cx<T> um = bU[m * kApplyPitch + lane];
#pragma unroll
      for (int u = 0; u < 4; ++u) { const T v = sVt[m * np + jb * 4 + u]; acc[u].re = fma(v, um.re, acc[u].re); acc[u].im = fma(v, um.im, acc[u].im); }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int j = jb * 4 + u; if (j < n) bT2[j * kApplyPitch + lane] = conj(bP[j * kApplyPitch + lane]) * acc[u]; }
  }
  __syncthreads();
  for (int e = tid; e < ns * n; e += 128) { const int sidx_ = e / n, k = e - sidx_ * n; Xbar[(size_t)b0 * n + e] = bT2[k * kApplyPitch + sidx_]; }
}

// ------------------------------------------------------------------ Givens-product unitary
// One warp per parameter set.  X (and its cotangent) live in shared memory column-major,
// lane = row: each rotation is a two-column update that touches only the lane's own row, so the
// whole product needs no synchronisation.  The reverse pass undoes the rotations (R is unitary:
// X_{k-1} = X_k R_k^H, no tape).  The product is ALWAYS accumulated in FP64 (W = double), also for
// complex64 I/O: N(N-1)/2 chained rotations in float32 miss the reference tests' 6-decimal
// unitarity bound at N >= 26, and the kernel is latency/HBM bound, so FP64 costs nothing.
// smem per warp: X (N*N) [+ G (N*N)] complex<double>, then trig (K x 4 doubles).
template <typename T, bool BWD>
__global__ void givens_kernel(const T* __restrict__ thetas, const T* __restrict__ phis, const T* __restrict__ omegas,
                              cx<T>* __restrict__ U, const cx<T>* __restrict__ Ubar, T* __restrict__ tbar,
                              T* __restrict__ pbar, T* __restrict__ obar, long long batch, int N, int warps_per_block) {
  using W = double;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = (long long)blockIdx.x * warps_per_block + warp;
  if (b >= batch) return;
  const int K = N * (N - 1) / 2;
  const size_t per_warp = (size_t)(BWD ? 2 : 1) * N * N * sizeof(cx<W>) + (size_t)4 * K * sizeof(W);
  unsigned char* base = smem_raw + warp * ((per_warp + 15) / 16 * 16);
  cx<W>* X = reinterpret_cast<cx<W>*>(base);
  cx<W>* G = X + (BWD ? N * N : 0);
  W* trig = reinterpret_cast<W*>(base + (size_t)(BWD ? 2 : 1) * N * N * sizeof(cx<W>));
  const T* th = thetas + (size_t)b * K;
  const T* ph = phis + (size_t)b * K;
  const T* om = omegas + (size_t)b * N;
  for (int k = lane; k < K; k += 32) {
    W s, c, sp, cp;
    sincos_((W)th[k], &s, &c);
    sincos_((W)ph[k], &sp, &cp);
    trig[4 * k + 0] = c; trig[4 * k + 1] = s; trig[4 * k + 2] = cp; trig[4 * k + 3] = sp;
  }
  for (int e = lane; e < N * N; e += 32) { const int c = e / N, r = e % N; X[e] = cx<W>(r == c ? W(1) : W(0), W(0)); }
  __syncwarp();
  // forward: X <- X R_k,  R_aa = e^{-i phi} cos, R_ab = e^{-i phi} sin, R_ba = -sin, R_bb = cos
  int k = 0;
  for (int i = 2; i <= N; ++i)
    for (int j = 1; j < i; ++j, ++k) {
      const int a = i - 1, bb = j - 1;
      const W c = trig[4 * k], s = trig[4 * k + 1];
      const cx<W> em(trig[4 * k + 2], -trig[4 * k + 3]);       // e^{-i phi}
      for (int r = lane; r < N; r += 32) {
        const cx<W> xa = X[a * N + r], xb = X[bb * N + r];
        const cx<W> ea = em * xa;
        X[a * N + r] = cx<W>(c * ea.re - s * xb.re, c * ea.im - s * xb.im);
        X[bb * N + r] = cx<W>(s * ea.re + c * xb.re, s * ea.im + c * xb.im);
      }
    }
  __syncwarp();
  if (!BWD) {
    for (int e = lane; e < N * N; e += 32) {
      const int r = e / N, c = e % N;
      const cx<W> u = expi<W>((W)om[r]) * X[c * N + r];
      U[(size_t)b * N * N + e] = cx<T>((T)u.re, (T)u.im);
    }
    return;
  }
  // cotangent of X and omegabar
  for (int r = lane; r < N; r += 32) {
    const cx<W> d = expi<W>((W)om[r]);
    W acc = W(0);
    for (int c = 0; c < N; ++c) {
      const cx<T> ubt = Ubar[(size_t)b * N * N + (size_t)r * N + c];
      const cx<W> ub((W)ubt.re, (W)ubt.im);
      const cx<W> u = d * X[c * N + r];
      acc += -(ub.re * u.im - ub.im * u.re);                   // -Im(conj(ub) u)
      G[c * N + r] = conj(d) * ub;
    }
    obar[(size_t)b * N + r] = (T)acc;
  }
  __syncwarp();
  for (int i = N; i >= 2; --i)
    for (int j = i - 1; j >= 1; --j) {
      --k;
      const int a = i - 1, bb = j - 1;
      const W c = trig[4 * k], s = trig[4 * k + 1];
      const cx<W> em(trig[4 * k + 2], -trig[4 * k + 3]);
      W tb = W(0), pb = W(0);
      for (int r = lane; r < N; r += 32) {
        // undo the rotation: X_{k-1} = X_k R^H
        const cx<W> ya = X[a * N + r], yb = X[bb * N + r];
        const cx<W> pa(c * ya.re + s * yb.re, c * ya.im + s * yb.im);
        const cx<W> xa = conj(em) * pa;                                     // conj(R_aa) ya + conj(R_ab) yb
        const cx<W> xb(-s * ya.re + c * yb.re, -s * ya.im + c * yb.im);     // conj(R_ba) ya + conj(R_bb) yb
        X[a * N + r] = xa; X[bb * N + r] = xb;
        const cx<W> ga = G[a * N + r], gb = G[bb * N + r];
        // d/dtheta: col a: -e^{-i phi} sin xa - cos xb ; col b: e^{-i phi} cos xa - sin xb
        const cx<W> ea = em * xa;
        const cx<W> dta(-s * ea.re - c * xb.re, -s * ea.im - c * xb.im);
        const cx<W> dtb(c * ea.re - s * xb.re, c * ea.im - s * xb.im);
        tb += ga.re * dta.re + ga.im * dta.im + gb.re * dtb.re + gb.im * dtb.im;
        // d/dphi: col a: -i e^{-i phi} cos xa ; col b: -i e^{-i phi} sin xa   (-i z = (z.im, -z.re))
        const cx<W> dpa(c * ea.im, -c * ea.re), dpb(s * ea.im, -s * ea.re);
        pb += ga.re * dpa.re + ga.im * dpa.im + gb.re * dpb.re + gb.im * dpb.im;
        // G_{k-1} = G_k R^H
        const cx<W> qa(c * ga.re + s * gb.re, c * ga.im + s * gb.im);
        G[a * N + r] = conj(em) * qa;
        G[bb * N + r] = cx<W>(-s * ga.re + c * gb.re, -s * ga.im + c * gb.im);
      }
      tb = warp_sum(tb); pb = warp_sum(pb);
      if (lane == 0) { tbar[(size_t)b * K + k] = (T)tb; pbar[(size_t)b * K + k] = (T)pb; }
    }
}

// Register-resident variant for N <= NT in {4, 8, 16, 32}: lane = (parameter set, row), 32 / NT sets per
// warp, and the lane's whole row of X (and of its cotangent) lives in REGISTERS -- the (i, j) loops are
// fully unrolled so every column index is static and a rotation is a dozen FP64 FMAs on registers
// instead of a shared-memory read-modify-write chain.  Only the sin/cos table goes through shared memory.
template <typename T, int NT, bool BWD>
__global__ void __launch_bounds__(128) givens_reg_kernel(const T* __restrict__ thetas, const T* __restrict__ phis,
                                                         const T* __restrict__ omegas, cx<T>* __restrict__ U,
                                                         const cx<T>* __restrict__ Ubar, T* __restrict__ tbar,
                                                         T* __restrict__ pbar, T* __restrict__ obar, long long batch, int N) {
  using W = double;
  constexpr int SPW = 32 / NT, KMAX = NT * (NT - 1) / 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane / NT, row = lane % NT;
  const long long b = ((long long)blockIdx.x * (blockDim.x >> 5) + warp) * SPW + grp;
  const bool live = b < batch;
  const int K = N * (N - 1) / 2;
  // per warp: trig [SPW][KMAX][4], then (BWD) bars [SPW][KMAX][2]
  W* trig = reinterpret_cast<W*>(smem_raw) + (size_t)warp * SPW * KMAX * (BWD ? 6 : 4) + (size_t)grp * KMAX * 4;
  W* bars = reinterpret_cast<W*>(smem_raw) + (size_t)warp * SPW * KMAX * 6 + (size_t)SPW * KMAX * 4 + (size_t)grp * KMAX * 2;
  if (live)
    for (int k = row; k < K; k += NT) {
      W s, c, sp, cp;
      sincos_((W)thetas[(size_t)b * K + k], &s, &c);
      sincos_((W)phis[(size_t)b * K + k], &sp, &cp);
      trig[4 * k + 0] = c; trig[4 * k + 1] = s; trig[4 * k + 2] = cp; trig[4 * k + 3] = sp;
    }
  __syncwarp();
  cx<W> x[NT];
#pragma unroll
  for (int c = 0; c < NT; ++c) x[c] = cx<W>(c == row ? W(1) : W(0), W(0));
  // forward: X <- X R_k,  R_aa = e^{-i phi} cos, R_ab = e^{-i phi} sin, R_ba = -sin, R_bb = cos
#pragma unroll
  for (int i = 2; i <= NT; ++i) {
    if (i <= N) {
#pragma unroll
      for (int j = 1; j < i; ++j) {
        constexpr int dummy = 0; (void)dummy;
        const int k = (i - 1) * (i - 2) / 2 + (j - 1);
        const double2 cs = *reinterpret_cast<const double2*>(&trig[4 * k]);
        const double2 ph = *reinterpret_cast<const double2*>(&trig[4 * k + 2]);
        const cx<W> em(ph.x, -ph.y);
        const cx<W> xa = x[i - 1], xb = x[j - 1];
        const cx<W> ea = em * xa;
        x[i - 1] = cx<W>(cs.x * ea.re - cs.y * xb.re, cs.x * ea.im - cs.y * xb.im);
        x[j - 1] = cx<W>(cs.y * ea.re + cs.x * xb.re, cs.y * ea.im + cs.x * xb.im);
      }
    }
  }
  const bool rowlive = live && row < N;
  const cx<W> d = rowlive ? expi<W>((W)omegas[(size_t)b * N + row]) : cx<W>(W(1), W(0));
  if (!BWD) {
    if (rowlive) {
      cx<T>* u = U + (size_t)b * N * N + (size_t)row * N;
#pragma unroll
      for (int c = 0; c < NT; ++c)
        if (c < N) { const cx<W> v = d * x[c]; u[c] = cx<T>((T)v.re, (T)v.im); }
    }
    return;
  }
  cx<W> g[NT];
  W ob = W(0);
#pragma unroll
  for (int c = 0; c < NT; ++c) {
    g[c] = cx<W>(W(0), W(0));
    if (rowlive && c < N) {
      const cx<T> ubt = Ubar[(size_t)b * N * N + (size_t)row * N + c];
      const cx<W> ub((W)ubt.re, (W)ubt.im);
      const cx<W> u = d * x[c];
      ob += -(ub.re * u.im - ub.im * u.re);                       // -Im(conj(ub) u)
      g[c] = conj(d) * ub;
    }
  }
  if (rowlive) obar[(size_t)b * N + row] = (T)ob;
#pragma unroll
  for (int i = NT; i >= 2; --i) {
    if (i <= N) {
#pragma unroll
      for (int j = i - 1; j >= 1; --j) {
        const int k = (i - 1) * (i - 2) / 2 + (j - 1);
        const double2 cs = *reinterpret_cast<const double2*>(&trig[4 * k]);
        const double2 ph = *reinterpret_cast<const double2*>(&trig[4 * k + 2]);
        const W c = cs.x, s = cs.y;
        const cx<W> em(ph.x, -ph.y);
        // undo the rotation: X_{k-1} = X_k R^H
        const cx<W> ya = x[i - 1], yb = x[j - 1];
        const cx<W> pa(c * ya.re + s * yb.re, c * ya.im + s * yb.im);
        const cx<W> xa = conj(em) * pa;
        const cx<W> xb(-s * ya.re + c * yb.re, -s * ya.im + c * yb.im);
        x[i - 1] = xa; x[j - 1] = xb;
        const cx<W> ga = g[i - 1], gb = g[j - 1];
        const cx<W> ea = em * xa;
        const cx<W> dta(-s * ea.re - c * xb.re, -s * ea.im - c * xb.im);
        const cx<W> dtb(c * ea.re - s * xb.re, c * ea.im - s * xb.im);
        W tb = ga.re * dta.re + ga.im * dta.im + gb.re * dtb.re + gb.im * dtb.im;
        const cx<W> dpa(c * ea.im, -c * ea.re), dpb(s * ea.im, -s * ea.re);
        W pb = ga.re * dpa.re + ga.im * dpa.im + gb.re * dpb.re + gb.im * dpb.im;
        const cx<W> qa(c * ga.re + s * gb.re, c * ga.im + s * gb.im);
        g[i - 1] = conj(em) * qa;
        g[j - 1] = cx<W>(-s * ga.re + c * gb.re, -s * ga.im + c * gb.im);
#pragma unroll
        for (int o = NT >> 1; o > 0; o >>= 1) { tb += __shfl_xor_sync(0xffffffffu, tb, o); pb += __shfl_xor_sync(0xffffffffu, pb, o); }
        if (row == 0) { bars[2 * k] = tb; bars[2 * k + 1] = pb; }
      }
    }
  }
  __syncwarp();
  if (live)
    for (int k = row; k < K; k += NT) { tbar[(size_t)b * K + k] = (T)bars[2 * k]; pbar[(size_t)b * K + k] = (T)bars[2 * k + 1]; }
}


// ------------------------------------------------------------------ fused Unitary-learning cost (config #1)
// examples/Unitary-Learning-qgrad.py:123-151 for batches of parameter sets sharing the M training pairs:
//   loss_b = 1 - (1/M) sum_k | Re < out_k | U_b | in_k > |,   U_b = Unitary(N)(thetas_b, phis_b, omegas_b),
// and d loss_b / d (thetas, phis, omegas) in the same pass.  Same lane = (set, row) layout as givens_reg_kernel: the
// lane builds its row of U in registers, applies it to every ket (kets live in shared memory, broadcast reads),
// the overlaps are reduced over the NT lanes of the set by shuffles, the cotangent row
// Ubar[r][:] = sum_k coef_k out_k[r] conj(in_k[:]), coef_k = -sign(Re o_k)/M, accumulates in registers and goes
// straight into the reverse Givens sweep -- U, U psi and Ubar never touch HBM.
template <typename T, int NT>
__global__ void __launch_bounds__(128) unitary_cost_kernel(const T* __restrict__ thetas, const T* __restrict__ phis,
                                                           const T* __restrict__ omegas, const cx<T>* __restrict__ kin,
                                                           const cx<T>* __restrict__ kout, int M, T* __restrict__ loss,
                                                           T* __restrict__ tbar, T* __restrict__ pbar, T* __restrict__ obar,
                                                           long long batch, int N) {
  using W = double;
  constexpr int SPW = 32 / NT, KMAX = NT * (NT - 1) / 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane / NT, row = lane % NT;
  const int nwarps = blockDim.x >> 5;
  const long long b = ((long long)blockIdx.x * nwarps + warp) * SPW + grp;
  const bool live = b < batch;
  const int K = N * (N - 1) / 2;
  // smem: kets in [M][NT], kets out [M][NT] (complex<double>, zero padded) | per warp: trig [SPW][KMAX][4], bars [SPW][KMAX][2]
  cx<W>* sin_ = reinterpret_cast<cx<W>*>(smem_raw);
  cx<W>* sout = sin_ + (size_t)M * NT;
  W* wbase = reinterpret_cast<W*>(sout + (size_t)M * NT) + (size_t)warp * SPW * KMAX * 6;
  W* trig = wbase + (size_t)grp * KMAX * 4;
  W* bars = wbase + (size_t)SPW * KMAX * 4 + (size_t)grp * KMAX * 2;
  for (int e = threadIdx.x; e < M * NT; e += blockDim.x) {
    const int k = e / NT, c = e - k * NT;
    cx<W> vi(W(0), W(0)), vo(W(0), W(0));
    if (c < N) { const cx<T> a = kin[(size_t)k * N + c], o = kout[(size_t)k * N + c]; vi = cx<W>((W)a.re, (W)a.im); vo = cx<W>((W)o.re, (W)o.im); }
    sin_[e] = vi; sout[e] = vo;
  }
  if (live)
    for (int k = row; k < K; k += NT) {
      W s, c, sp, cp;
      sincos_((W)thetas[(size_t)b * K + k], &s, &c);
      sincos_((W)phis[(size_t)b * K + k], &sp, &cp);
      trig[4 * k + 0] = c; trig[4 * k + 1] = s; trig[4 * k + 2] = cp; trig[4 * k + 3] = sp;
    }
  __syncthreads();
  cx<W> x[NT];
#pragma unroll
  for (int c = 0; c < NT; ++c) x[c] = cx<W>(c == row ? W(1) : W(0), W(0));
#pragma unroll
  for (int i = 2; i <= NT; ++i) {
    if (i <= N) {
#pragma unroll
      for (int j = 1; j < i; ++j) {
        const int k = (i - 1) * (i - 2) / 2 + (j - 1);
        const double2 cs = *reinterpret_cast<const double2*>(&trig[4 * k]);
        const double2 ph = *reinterpret_cast<const double2*>(&trig[4 * k + 2]);
        const cx<W> em(ph.x, -ph.y);
        const cx<W> xa = x[i - 1], xb = x[j - 1];
        const cx<W> ea = em * xa;
        x[i - 1] = cx<W>(cs.x * ea.re - cs.y * xb.re, cs.x * ea.im - cs.y * xb.im);
        x[j - 1] = cx<W>(cs.y * ea.re + cs.x * xb.re, cs.y * ea.im + cs.x * xb.im);
      }
    }
  }
  const bool rowlive = live && row < N;
  const cx<W> d = rowlive ? expi<W>((W)omegas[(size_t)b * N + row]) : cx<W>(W(1), W(0));
  // ---- the kets: overlaps, loss, and the cotangent row of U
  cx<W> ub[NT];
  cx<W> (&u)[NT] = x;                                                // x becomes the row of U in place (undone below: |d| = 1)
#pragma unroll
  for (int c = 0; c < NT; ++c) { u[c] = rowlive ? d * x[c] : cx<W>(W(0), W(0)); ub[c] = cx<W>(W(0), W(0)); }
  W acc_loss = W(0);
  const W inv_m = W(1) / (W)M;
  for (int k = 0; k < M; ++k) {
    cx<W> p(W(0), W(0));
#pragma unroll
    for (int c = 0; c < NT; ++c) cfma(p, u[c], sin_[k * NT + c]);
    const cx<W> o_r = sout[k * NT + row];
    cx<W> ov(W(0), W(0));
    cfma_conj(ov, o_r, p);                                          // conj(out_k[r]) pred_k[r]
    W ore = ov.re;
#pragma unroll
    for (int off = 1; off < NT; off <<= 1) ore += __shfl_xor_sync(0xffffffffu, ore, off);
    acc_loss += fabs(ore);
    const W coef = ore > W(0) ? -inv_m : (ore < W(0) ? inv_m : W(0));
    const cx<W> go(coef * o_r.re, coef * o_r.im);                   // cotangent of pred_k[r]
#pragma unroll
    for (int c = 0; c < NT; ++c) cfma(ub[c], go, conj(sin_[k * NT + c]));
  }
  if (live && row == 0) loss[b] = (T)(W(1) - acc_loss * inv_m);
  if (!tbar) return;
  // ---- reverse Givens sweep (as givens_reg_kernel<BWD>)
  cx<W> g[NT];
  W ob = W(0);
#pragma unroll
  for (int c = 0; c < NT; ++c) {
    g[c] = cx<W>(W(0), W(0));
    if (rowlive && c < N) {
      ob += -(ub[c].re * u[c].im - ub[c].im * u[c].re);             // -Im(conj(ub) u)
      g[c] = conj(d) * ub[c];
    }
    x[c] = conj(d) * u[c];                                          // back to the row of X
  }
  if (rowlive) obar[(size_t)b * N + row] = (T)ob;
#pragma unroll
  for (int i = NT; i >= 2; --i) {
    if (i <= N) {
#pragma unroll
      for (int j = i - 1; j >= 1; --j) {
        const int k = (i - 1) * (i - 2) / 2 + (j - 1);
        const double2 cs = *reinterpret_cast<const double2*>(&trig[4 * k]);
        const double2 ph = *reinterpret_cast<const double2*>(&trig[4 * k + 2]);
        const W c = cs.x, s = cs.y;
        const cx<W> em(ph.x, -ph.y);
        const cx<W> ya = x[i - 1], yb = x[j - 1];
        const cx<W> pa(c * ya.re + s * yb.re, c * ya.im + s * yb.im);
        const cx<W> xa = conj(em) * pa;
        const cx<W> xb(-s * ya.re + c * yb.re, -s * ya.im + c * yb.im);
        x[i - 1] = xa; x[j - 1] = xb;
        const cx<W> ga = g[i - 1], gb = g[j - 1];
        const cx<W> ea = em * xa;
        const cx<W> dta(-s * ea.re - c * xb.re, -s * ea.im - c * xb.im);
        const cx<W> dtb(c * ea.re - s * xb.re, c * ea.im - s * xb.im);
        W tb = ga.re * dta.re + ga.im * dta.im + gb.re * dtb.re + gb.im * dtb.im;
        const cx<W> dpa(c * ea.im, -c * ea.re), dpb(s * ea.im, -s * ea.re);
        W pb = ga.re * dpa.re + ga.im * dpa.im + gb.re * dpb.re + gb.im * dpb.im;
        const cx<W> qa(c * ga.re + s * gb.re, c * ga.im + s * gb.im);
        g[i - 1] = conj(em) * qa;
        g[j - 1] = cx<W>(-s * ga.re + c * gb.re, -s * ga.im + c * gb.im);
#pragma unroll
        for (int o = NT >> 1; o > 0; o >>= 1) { tb += __shfl_xor_sync(0xffffffffu, tb, o); pb += __shfl_xor_sync(0xffffffffu, pb, o); }
        if (row == 0) { bars[2 * k] = tb; bars[2 * k + 1] = pb; }
      }
    }
  }
  __syncwarp();
  if (live)
    for (int k = row; k < K; k += NT) { tbar[(size_t)b * K + k] = (T)bars[2 * k]; pbar[(size_t)b * K + k] = (T)bars[2 * k + 1]; }
}

// ------------------------------------------------------------------ _make_rot
// grid (ceil(N*N / 256), ceil(batch / kRotMatsPerBlock)): the block first evaluates sin/cos(theta) and e^{i phi}
// of its kRotMatsPerBlock matrices once, one matrix per thread, into shared memory -- done inside the store loop the
// four special positions would make every warp run the (divergent, ~100-instruction) sincos path per matrix.
// Then a thread owns ONE position (r, c), decoded once, and walks the matrices: the loop is a select and a store.
constexpr int kRotMatsPerBlock = 64;
template <typename T>
__global__ void make_rot_fwd_kernel(const T* __restrict__ params, cx<T>* __restrict__ R, long long batch, int N, int i, int j) {
  __shared__ T ssin[kRotMatsPerBlock], scos[kRotMatsPerBlock];
  __shared__ cx<T> sep[kRotMatsPerBlock];
  const long long b0 = (long long)blockIdx.y * kRotMatsPerBlock;
  const long long b1 = b0 + kRotMatsPerBlock < batch ? b0 + kRotMatsPerBlock : batch;
  if (threadIdx.x < kRotMatsPerBlock && b0 + threadIdx.x < b1) {
    const long long b = b0 + threadIdx.x;
    T s, co; sincos_(params[2 * b], &s, &co);
    ssin[threadIdx.x] = s; scos[threadIdx.x] = co; sep[threadIdx.x] = expi<T>(params[2 * b + 1]);
  }
  __syncthreads();
  // small matrices: `span` = next power of two >= N*N positions per matrix, 256 / span matrices side by side
  int span = 1;
  while (span < N * N && span < 256) span <<= 1;
  const int lanes = 256 / span, sub = threadIdx.x / span;
  const int rc = blockIdx.x * span + (threadIdx.x - sub * span);
  if (rc >= N * N) return;
  const int r = rc / N, c = rc - r * N;
  // four successive writes in the reference, later ones win (matters only for i == j)
  const int kind = (r == j && c == j) ? 4 : (r == j && c == i) ? 3 : (r == i && c == j) ? 2 : (r == i && c == i) ? 1 : 0;
  const cx<T> plain(r == c ? T(1) : T(0), T(0));
  for (long long b = b0 + sub; b < b1; b += lanes) {
    const int q = (int)(b - b0);
    cx<T> v = plain;
    if (kind == 1) v = scos[q] * sep[q];
    else if (kind == 2) v = -(ssin[q] * sep[q]);
    else if (kind == 3) v = cx<T>(ssin[q], T(0));
    else if (kind == 4) v = cx<T>(scos[q], T(0));
    R[(size_t)b * N * N + rc] = v;
  }
}
template <typename T>
__global__ void make_rot_bwd_kernel(const T* __restrict__ params, const cx<T>* __restrict__ Rbar, T* __restrict__ pbar,
                                    long long batch, int N, int i, int j) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  T s, co; sincos_(params[2 * b], &s, &co);
  const cx<T> ep = expi<T>(params[2 * b + 1]);
  const cx<T>* g = Rbar + (size_t)b * N * N;
  auto rdot = [](cx<T> gg, cx<T> d) { return gg.re * d.re + gg.im * d.im; };   // Re(conj(g) d)
  T tb = T(0), pb = T(0);
  if (i != j) {
    const cx<T> iep(-ep.im, ep.re);                                            // i e^{i phi}
    tb += rdot(g[(size_t)i * N + i], -(s * ep)) + rdot(g[(size_t)i * N + j], -(co * ep)) + rdot(g[(size_t)j * N + i], cx<T>(co, T(0)));
    pb += rdot(g[(size_t)i * N + i], co * iep) + rdot(g[(size_t)i * N + j], -(s * iep));
  }
  tb += rdot(g[(size_t)j * N + j], cx<T>(-s, T(0)));
  pbar[2 * b] = tb; pbar[2 * b + 1] = pb;
}

// ------------------------------------------------------------------ ZYZ rotation
template <typename T> struct Zyz {
  cx<T> r00, r01, r10, r11, ep, em; T c, s;
  __device__ Zyz(T phi, T theta, T omega) {
    sincos_(theta * T(0.5), &s, &c);
    ep = expi<T>(T(0.5) * (phi + omega));      // e^{+i p}
    em = expi<T>(T(0.5) * (phi - omega));      // e^{+i m}
    r00 = c * conj(ep); r01 = -(s * em); r10 = s * conj(em); r11 = c * ep;
  }
};
template <typename T> __device__ __forceinline__ cx<T> times_i(cx<T> z) { return cx<T>(-z.im, z.re); }

template <typename T>
__global__ void rot_zyz_fwd_kernel(const T* __restrict__ ang, cx<T>* __restrict__ R, long long batch) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const Zyz<T> z(ang[3 * b], ang[3 * b + 1], ang[3 * b + 2]);
  cx<T>* r = R + 4 * b;
  r[0] = z.r00; r[1] = z.r01; r[2] = z.r10; r[3] = z.r11;
}
template <typename T>
__global__ void rot_zyz_bwd_kernel(const T* __restrict__ ang, const cx<T>* __restrict__ Rbar, T* __restrict__ abar, long long batch) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const Zyz<T> z(ang[3 * b], ang[3 * b + 1], ang[3 * b + 2]);
  const cx<T>* g = Rbar + 4 * b;
  auto rdot = [](cx<T> gg, cx<T> d) { return gg.re * d.re + gg.im * d.im; };
  const T h = T(0.5);
  const cx<T> i00 = times_i(z.r00), i01 = times_i(z.r01), i10 = times_i(z.r10), i11 = times_i(z.r11);
  // d/dphi: (-i/2, +i/2, -i/2, +i/2) ; d/domega: (-i/2, -i/2, +i/2, +i/2)
  abar[3 * b + 0] = h * (-rdot(g[0], i00) + rdot(g[1], i01) - rdot(g[2], i10) + rdot(g[3], i11));
  abar[3 * b + 2] = h * (-rdot(g[0], i00) - rdot(g[1], i01) + rdot(g[2], i10) + rdot(g[3], i11));
  // d/dtheta: (-e^{-ip} s, -e^{im} c, e^{-im} c, -e^{ip} s) / 2
  abar[3 * b + 1] = h * (rdot(g[0], -(z.s * conj(z.ep))) + rdot(g[1], -(z.c * z.em)) + rdot(g[2], z.c * conj(z.em)) + rdot(g[3], -(z.s * z.ep)));
}
// F = |(R ket)_0|^2 and dF/d(phi, theta, omega)
template <typename T>
__global__ void rot_fidelity_kernel(const T* __restrict__ ang, const cx<T>* __restrict__ ket, int ket_batched,
                                    T* __restrict__ fid, T* __restrict__ grad, long long batch) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const Zyz<T> z(ang[3 * b], ang[3 * b + 1], ang[3 * b + 2]);
  const cx<T> k0 = ket[ket_batched ? 2 * b : 0], k1 = ket[ket_batched ? 2 * b + 1 : 1];
  const cx<T> t0 = z.r00 * k0, t1 = z.r01 * k1;
  const cx<T> v = t0 + t1;
  fid[b] = v.re * v.re + v.im * v.im;
  if (grad) {
    auto rdot = [](cx<T> gg, cx<T> d) { return gg.re * d.re + gg.im * d.im; };
    const cx<T> it0 = times_i(t0), it1 = times_i(t1);
    const cx<T> dphi(T(0.5) * (-it0.re + it1.re), T(0.5) * (-it0.im + it1.im));
    const cx<T> dom(T(0.5) * (-it0.re - it1.re), T(0.5) * (-it0.im - it1.im));
    const cx<T> dth = T(0.5) * ((-(z.s * conj(z.ep))) * k0 + (-(z.c * z.em)) * k1);
    grad[3 * b + 0] = T(2) * rdot(v, dphi);
    grad[3 * b + 1] = T(2) * rdot(v, dth);
    grad[3 * b + 2] = T(2) * rdot(v, dom);
  }
}

// ------------------------------------------------------------------ host launchers
template <typename T>
static int displace(bool bwd, const void* V, const void* lam, const void* alpha, void* D, const void* Dbar, void* abar,
                    int64_t batch, int n, cudaStream_t st) {
  {
    // register-tiled kernel: up to n = 44 (121 threads per alpha)
    const int np = (n + 3) / 4 * 4, tpa = (np / 4) * (np / 4);
    if (tpa <= 128) {
      const int apb = 128 / tpa;
      const size_t smem = (size_t)np * np * sizeof(T) + (size_t)apb * 3 * np * sizeof(cx<T>) + (size_t)128 * 2 * sizeof(T) + 16;
      auto kf = displace_tiled_kernel<T, false>; auto kb = displace_tiled_kernel<T, true>;
      if (smem > 48 * 1024) QG_RETURN_IF_CUDA(cudaFuncSetAttribute(bwd ? kb : kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      const unsigned grid = (unsigned)((batch + apb - 1) / apb);
      if (bwd) kb<<<grid, 128, smem, st>>>((const T*)V, (const T*)lam, (const cx<T>*)alpha, nullptr, (const cx<T>*)Dbar, (cx<T>*)abar, batch, n, np, apb);
      else kf<<<grid, 128, smem, st>>>((const T*)V, (const T*)lam, (const cx<T>*)alpha, (cx<T>*)D, nullptr, nullptr, batch, n, np, apb);
      QG_CHECK_LAUNCH();
      return QG_OK;
    }
  }
  const size_t smem = ((size_t)n * n + 1) * sizeof(T) + (size_t)3 * n * sizeof(cx<T>) + 32 * sizeof(T) + 16;
  if (smem > 227 * 1024) return QG_ERR_DIM;
  auto kf = displace_kernel<T, false>; auto kb = displace_kernel<T, true>;
  if (smem > 48 * 1024) QG_RETURN_IF_CUDA(cudaFuncSetAttribute(bwd ? kb : kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int threads = n * n >= 256 ? 256 : (n * n >= 64 ? 128 : 32);
  if (bwd) kb<<<(unsigned)batch, threads, smem, st>>>((const T*)V, (const T*)lam, (const cx<T>*)alpha, nullptr, (const cx<T>*)Dbar, (cx<T>*)abar, n);
  else kf<<<(unsigned)batch, threads, smem, st>>>((const T*)V, (const T*)lam, (const cx<T>*)alpha, (cx<T>*)D, nullptr, nullptr, n);
  QG_CHECK_LAUNCH();
  return QG_OK;
}
template <typename T>
static int displace_apply(bool bwd, const void* V, const void* lam, const void* alpha, const void* x, void* y,
                          const void* ybar, void* abar, void* xbar, int64_t batch, int n, cudaStream_t st) {
  {
    // 32 kets per CTA, lane = sample (falls back to one CTA per ket when the buffers outgrow shared memory)
    const int np = (n + 3) / 4 * 4;
    const size_t smem_b = (size_t)(2 * n * np + np + 64 + 256) * sizeof(T) + (size_t)(bwd ? 6 : 3) * np * kApplyPitch * sizeof(cx<T>) + 16;
    if (smem_b <= 200 * 1024 && batch >= 32) {
      auto kf = displace_apply_batched_kernel<T, false>; auto kb = displace_apply_batched_kernel<T, true>;
      if (smem_b > 48 * 1024) QG_RETURN_IF_CUDA(cudaFuncSetAttribute(bwd ? kb : kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
      const unsigned grid = (unsigned)((batch + 31) / 32);
      if (bwd) kb<<<grid, 128, smem_b, st>>>((const T*)V, (const T*)lam, (const cx<T>*)alpha, (const cx<T>*)x, nullptr, (const cx<T>*)ybar, (cx<T>*)abar, (cx<T>*)xbar, batch, n, np);
      else kf<<<grid, 128, smem_b, st>>>((const T*)V, (const T*)lam, (const cx<T>*)alpha, (const cx<T>*)x, (cx<T>*)y, nullptr, nullptr, nullptr, batch, n, np);
      QG_CHECK_LAUNCH();
      return QG_OK;
    }
  }
  const size_t smem = (size_t)6 * n * sizeof(cx<T>) + 32 * sizeof(T) + 16;
  if (smem > 48 * 1024) return QG_ERR_DIM;
  const int threads = n >= 64 ? 128 : (n > 32 ? 64 : 32);
  if (bwd) displace_apply_kernel<T, true><<<(unsigned)batch, threads, smem, st>>>((const T*)V, (const T*)lam, (const cx<T>*)alpha, (const cx<T>*)x, nullptr, (const cx<T>*)ybar, (cx<T>*)abar, (cx<T>*)xbar, n);
  else displace_apply_kernel<T, false><<<(unsigned)batch, threads, smem, st>>>((const T*)V, (const T*)lam, (const cx<T>*)alpha, (const cx<T>*)x, (cx<T>*)y, nullptr, nullptr, nullptr, n);
  QG_CHECK_LAUNCH();
  return QG_OK;
}
template <typename T>
static int givens(bool bwd, const void* th, const void* ph, const void* om, void* U, const void* Ubar, void* tb, void* pb,
                  void* ob, int64_t batch, int N, cudaStream_t st) {
  if (N < 1) return QG_ERR_DIM;
  // register-resident kernels: N <= 16 (both passes), N <= 32 (forward)
  if (N <= 16 || (N <= 32 && !bwd)) {
    const int NT = N <= 4 ? 4 : (N <= 8 ? 8 : (N <= 16 ? 16 : 32));
    const int spw = 32 / NT, kmax = NT * (NT - 1) / 2, warps = 4;
    const size_t smem = (size_t)warps * spw * kmax * (bwd ? 6 : 4) * sizeof(double);
    const unsigned grid = (unsigned)((batch + (int64_t)warps * spw - 1) / ((int64_t)warps * spw));
#define QG_GIVENS_REG(NTV)                                                                                                   \
    do {                                                                                                                     \
      auto kf = givens_reg_kernel<T, NTV, false>; auto kb = givens_reg_kernel<T, NTV, true>;                                 \
      if (smem > 48 * 1024) QG_RETURN_IF_CUDA(cudaFuncSetAttribute(bwd ? kb : kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      if (bwd) kb<<<grid, warps * 32, smem, st>>>((const T*)th, (const T*)ph, (const T*)om, nullptr, (const cx<T>*)Ubar, (T*)tb, (T*)pb, (T*)ob, batch, N); \
      else kf<<<grid, warps * 32, smem, st>>>((const T*)th, (const T*)ph, (const T*)om, (cx<T>*)U, nullptr, nullptr, nullptr, nullptr, batch, N); \
    } while (0)
    if (NT == 4) QG_GIVENS_REG(4); else if (NT == 8) QG_GIVENS_REG(8); else if (NT == 16) QG_GIVENS_REG(16); else QG_GIVENS_REG(32);
#undef QG_GIVENS_REG
    QG_CHECK_LAUNCH();
    return QG_OK;
  }
  const int K = N * (N - 1) / 2;
  size_t per_warp = (size_t)(bwd ? 2 : 1) * N * N * sizeof(cx<double>) + (size_t)4 * K * sizeof(double);
  per_warp = (per_warp + 15) / 16 * 16;
  if (per_warp > 200 * 1024) return QG_ERR_DIM;
  int wpb = (int)((64 * 1024) / per_warp);
  if (wpb > 4) wpb = 4;
  if (wpb < 1) wpb = 1;
  const size_t smem = per_warp * wpb;
  auto kf = givens_kernel<T, false>; auto kb = givens_kernel<T, true>;
  if (smem > 48 * 1024) QG_RETURN_IF_CUDA(cudaFuncSetAttribute(bwd ? kb : kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned grid = (unsigned)((batch + wpb - 1) / wpb);
  if (bwd) kb<<<grid, wpb * 32, smem, st>>>((const T*)th, (const T*)ph, (const T*)om, nullptr, (const cx<T>*)Ubar, (T*)tb, (T*)pb, (T*)ob, batch, N, wpb);
  else kf<<<grid, wpb * 32, smem, st>>>((const T*)th, (const T*)ph, (const T*)om, (cx<T>*)U, nullptr, nullptr, nullptr, nullptr, batch, N, wpb);
  QG_CHECK_LAUNCH();
  return QG_OK;
}

template <typename T>
static int unitary_cost(const void* th, const void* ph, const void* om, const void* kin, const void* kout, int M, void* loss,
                        void* tb, void* pb, void* ob, int64_t batch, int N, cudaStream_t st) {
  if (N < 1 || N > 16 || M < 1) return QG_ERR_DIM;
  const int NT = N <= 4 ? 4 : (N <= 8 ? 8 : 16);
  const int spw = 32 / NT, kmax = NT * (NT - 1) / 2, warps = 4;
  const size_t smem = (size_t)2 * M * NT * sizeof(cx<double>) + (size_t)warps * spw * kmax * 6 * sizeof(double);
  if (smem > 200 * 1024) return QG_ERR_DIM;
  const unsigned grid = (unsigned)((batch + (int64_t)warps * spw - 1) / ((int64_t)warps * spw));
#define QG_UCOST(NTV)                                                                                                          \
  do {                                                                                                                         \
    auto kf = unitary_cost_kernel<T, NTV>;                                                                                     \
    if (smem > 48 * 1024) QG_RETURN_IF_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kf<<<grid, warps * 32, smem, st>>>((const T*)th, (const T*)ph, (const T*)om, (const cx<T>*)kin, (const cx<T>*)kout, M,     \
                                      (T*)loss, (T*)tb, (T*)pb, (T*)ob, batch, N);                                             \
  } while (0)
  if (NT == 4) QG_UCOST(4); else if (NT == 8) QG_UCOST(8); else QG_UCOST(16);
#undef QG_UCOST
  QG_CHECK_LAUNCH();
  return QG_OK;
}

template <typename T>
static int misc(int op, const void* p0, const void* p1, void* o0, void* o1, int64_t batch, int N, int i, int j, int flag,
                cudaStream_t st) {
  const int threads = 256;
  const unsigned gb = (unsigned)((batch + threads - 1) / threads);
  switch (op) {
    case 0: {
      const long long gy = (batch + kRotMatsPerBlock - 1) / kRotMatsPerBlock;
      if (gy > 65535) {     // more than 4M matrices: in slices of 65535 * kRotMatsPerBlock
        const long long per = 65535LL * kRotMatsPerBlock;
        for (long long b0 = 0; b0 < batch; b0 += per) {
          const long long nb = batch - b0 < per ? batch - b0 : per;
          make_rot_fwd_kernel<T><<<dim3(N * N >= 256 ? (N * N + 255) / 256 : 1, (unsigned)((nb + kRotMatsPerBlock - 1) / kRotMatsPerBlock)), threads, 0, st>>>(
              (const T*)p0 + 2 * b0, (cx<T>*)o0 + (size_t)b0 * N * N, nb, N, i, j);
        }
      } else {
        make_rot_fwd_kernel<T><<<dim3(N * N >= 256 ? (N * N + 255) / 256 : 1, (unsigned)gy), threads, 0, st>>>((const T*)p0, (cx<T>*)o0, batch, N, i, j);
      }
      break;
    }
    case 1: make_rot_bwd_kernel<T><<<gb, threads, 0, st>>>((const T*)p0, (const cx<T>*)p1, (T*)o0, batch, N, i, j); break;
    case 2: rot_zyz_fwd_kernel<T><<<gb, threads, 0, st>>>((const T*)p0, (cx<T>*)o0, batch); break;
    case 3: rot_zyz_bwd_kernel<T><<<gb, threads, 0, st>>>((const T*)p0, (const cx<T>*)p1, (T*)o0, batch); break;
    case 4: rot_fidelity_kernel<T><<<gb, threads, 0, st>>>((const T*)p0, (const cx<T>*)p1, flag, (T*)o0, (T*)o1, batch); break;
    default: return QG_ERR_ARG;
  }
  QG_CHECK_LAUNCH();
  return QG_OK;
}

}  // namespace build

int builder_displace(int dtype, bool bwd, const void* V, const void* lam, const void* alpha, void* D, const void* Dbar,
                     void* abar, int64_t batch, int n, cudaStream_t st) {
  return dtype == QG_C128 ? build::displace<double>(bwd, V, lam, alpha, D, Dbar, abar, batch, n, st)
                          : build::displace<float>(bwd, V, lam, alpha, D, Dbar, abar, batch, n, st);
}
int builder_displace_apply(int dtype, bool bwd, const void* V, const void* lam, const void* alpha, const void* x, void* y,
                           const void* ybar, void* abar, void* xbar, int64_t batch, int n, cudaStream_t st) {
  return dtype == QG_C128 ? build::displace_apply<double>(bwd, V, lam, alpha, x, y, ybar, abar, xbar, batch, n, st)
                          : build::displace_apply<float>(bwd, V, lam, alpha, x, y, ybar, abar, xbar, batch, n, st);
}
int builder_givens(int dtype, bool bwd, const void* th, const void* ph, const void* om, void* U, const void* Ubar, void* tb,
                   void* pb, void* ob, int64_t batch, int N, cudaStream_t st) {
  return dtype == QG_C128 ? build::givens<double>(bwd, th, ph, om, U, Ubar, tb, pb, ob, batch, N, st)
                          : build::givens<float>(bwd, th, ph, om, U, Ubar, tb, pb, ob, batch, N, st);
}
int builder_unitary_cost(int dtype, const void* th, const void* ph, const void* om, const void* kin, const void* kout, int M,
                         void* loss, void* tb, void* pb, void* ob, int64_t batch, int N, cudaStream_t st) {
  return dtype == QG_C128 ? build::unitary_cost<double>(th, ph, om, kin, kout, M, loss, tb, pb, ob, batch, N, st)
                          : build::unitary_cost<float>(th, ph, om, kin, kout, M, loss, tb, pb, ob, batch, N, st);
}
int builder_misc(int dtype, int op, const void* p0, const void* p1, void* o0, void* o1, int64_t batch, int N, int i, int j,
                 int flag, cudaStream_t st) {
  return dtype == QG_C128 ? build::misc<double>(op, p0, p1, o0, o1, batch, N, i, j, flag, st)
                          : build::misc<float>(op, p0, p1, o0, o1, batch, N, i, j, flag, st);
}

}  // namespace qg

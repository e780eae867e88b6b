// Batched Hermitian eigen-problems on the device for density-matrix workloads: the
// eigendecomposition itself, the principal square root of a positive semi-definite matrix, the TRUE mixed-state
// (Uhlmann-Jozsa) fidelity F = (tr sqrt(sqrt(rho) sigma sqrt(rho)))^2 -- the quantity the reference's `_fidelity_dm`
// (qgrad/qgrad_qutip.py:50-64, a diagonal-only surrogate; `sqrtm` is imported at :7 and never used) stands in for --
// and the device halves of `isherm` / `isdm` (:357-396): exact Hermiticity, trace, smallest eigenvalue.
//
// One CTA per matrix, everything in shared memory (n <= 64: cavity truncations 10-40 and few-qubit registers are what
// qgrad's examples hold as density matrices).  Eigen-solver: cyclic two-sided Jacobi with complex rotations in the
// round-robin (tournament) ordering -- floor(m/2) disjoint rotations per step, m - 1 steps per sweep (m = n rounded up
// to even; the odd index out sits the step out), each step three barrier-separated phases:
//   A. one thread per pair (p, q): from a_pp, a_qq, a_pq = |g| e^{i phi}:  tau = (a_qq - a_pp) / 2|g|,
//      t = sgn(tau) / (|tau| + sqrt(1 + tau^2)),  c = 1 / sqrt(1 + t^2),  s = t c,   J = [[c, s e^{i phi}], [-s e^{-i phi}, c]]
//   B. columns:  A <- A J,  V <- V J          (pairs x n independent two-element updates)
//   C. rows:     A <- J^H A
// Converges quadratically: 6-8 sweeps in complex128 for n = 40; stops when the directly summed off-diagonal mass is
// below round-off of the Frobenius norm.  Jacobi's accuracy on positive semi-definite matrices is what the square
// roots below want: small eigenvalues come out with small ABSOLUTE error ~ u ||A||, so sqrt(max(lambda, 0)) is safe.
// Row pitch n | 1 keeps the column phase off a single bank.
#include "common.cuh"

namespace qg {
namespace spectral {

constexpr int NT = 256, NMAX = 64;
enum { MODE_EIGH = 0, MODE_SQRTM = 1, MODE_FIDELITY = 2 };

// The sweeps always run in FLOAT64, whatever the storage type: complex64 inputs are widened on load and narrowed on
// store (float32 rotations accumulate to 1.6e-5 in V after ~8 sweeps x 63 steps at n = 64 -- above the 1e-5 bar; this
// is not a throughput path, and the float64 shared-memory footprint still fits: 3 x 64 x 65 x 16 B = 200 KB).
template <typename T> struct Eps;
template <> struct Eps<double> { static constexpr double u = 1.2e-16; static constexpr int sweeps = 18; };
// round-off floor of the eigenvalues, relative to the largest: the solver's own (8 n u in float64) or, for complex64
// STORAGE, what rounding the inputs to float32 does to a null eigenvalue (~ u32 ||A||_F <= u32 sqrt(n) lambda_max)
template <typename TIO> __device__ __forceinline__ double noise_floor(int n) {
  const double own = 8.0 * n * Eps<double>::u;
  const double io = sizeof(TIO) == 4 ? 4.0 * sqrt((double)n) * 6e-8 : 0.0;
  return own > io ? own : io;
}
// eigenvalues of a positive semi-definite matrix that are zero up to round-off (noise_floor below).
// Taking their square roots would turn u-sized noise into sqrt(u)-sized contributions (the pure-state case: n - 1 such
// eigenvalues); they are clamped to zero instead.  A genuine eigenvalue below the threshold loses at most sqrt of it,
// which is what the noise would have cost anyway.

template <typename T>
__device__ __forceinline__ T block_sum(T v, T* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  T tot = T(0);
#pragma unroll
  for (int w = 0; w < NT / 32; ++w) tot += red[w];
  return tot;
}

// In-place Jacobi diagonalisation of the Hermitian A (n x n, pitch ld) in shared memory; V (may be null) receives the
// accumulated rotations (must hold the identity on entry).  On exit the eigenvalues are Re diag(A).
template <typename T>
__device__ void jacobi(cx<T>* A, cx<T>* V, int n, int ld, T* rot_c, cx<T>* rot_s, int* rot_p, int* rot_q, T* red, int* flag) {
  const int m = n + (n & 1), half = m / 2;
  for (int sweep = 0; sweep < Eps<T>::sweeps; ++sweep) {
    T off = T(0), tot = T(0);
    for (int e = threadIdx.x; e < n * n; e += NT) {
      const int r = e / n, c = e - r * n;
      const T v = cabs2_(A[r * ld + c]);
      tot += v;
      if (r != c) off += v;
    }
    off = block_sum(off, red);
    tot = block_sum(tot, red);
    if (threadIdx.x == 0) *flag = (off <= Eps<T>::u * Eps<T>::u * tot * (T)n || !(tot == tot)) ? 1 : 0;   // converged, or NaN: give up
    __syncthreads();
    if (*flag) break;
    for (int step = 0; step < m - 1; ++step) {
      // ---- A: rotation parameters, one thread per pair
      if (threadIdx.x < half) {
        const int k = threadIdx.x;
        int p = k == 0 ? m - 1 : (step + k) % (m - 1);
        int q = k == 0 ? step : (step - k + (m - 1)) % (m - 1);
        if (p > q) { const int t_ = p; p = q; q = t_; }
        T c = T(1);
        cx<T> s(T(0), T(0));
        if (q < n) {
          const cx<T> g = A[p * ld + q];
          const T ag = cabs_(g);
          if (ag > T(0)) {
            const T tau = (A[q * ld + q].re - A[p * ld + p].re) / (T(2) * ag);
            const T t = tau == T(0) ? T(1) : copysign(T(1), tau) / (fabs(tau) + sqrt(fma(tau, tau, T(1))));
            c = T(1) / sqrt(fma(t, t, T(1)));
            const T sn = t * c / ag;
            s = cx<T>(sn * g.re, sn * g.im);                                          // s e^{i phi}
          }
        } else {
          p = q = -1;                                                                 // the odd index out
        }
        rot_c[k] = c; rot_s[k] = s; rot_p[k] = p; rot_q[k] = q;
      }
      __syncthreads();
      // ---- B: columns of A and V.  col_p' = c col_p - conj(s) col_q ;  col_q' = s col_p + c col_q
      for (int it = threadIdx.x; it < half * n; it += NT) {
        const int k = it / n, i = it - k * n;
        const int p = rot_p[k], q = rot_q[k];
        if (p < 0) continue;
        const T c = rot_c[k];
        const cx<T> s = rot_s[k];
        {
          const cx<T> xp = A[i * ld + p], xq = A[i * ld + q];
          A[i * ld + p] = c * xp - conj(s) * xq;
          A[i * ld + q] = s * xp + c * xq;
        }
        if (V) {
          const cx<T> xp = V[i * ld + p], xq = V[i * ld + q];
          V[i * ld + p] = c * xp - conj(s) * xq;
          V[i * ld + q] = s * xp + c * xq;
        }
      }
      __syncthreads();
      // ---- C: rows of A.  row_p' = c row_p - s row_q ;  row_q' = conj(s) row_p + c row_q
      for (int it = threadIdx.x; it < half * n; it += NT) {
        const int k = it / n, j = it - k * n;
        const int p = rot_p[k], q = rot_q[k];
        if (p < 0) continue;
        const T c = rot_c[k];
        const cx<T> s = rot_s[k];
        const cx<T> xp = A[p * ld + j], xq = A[q * ld + j];
        cx<T> np_ = c * xp - s * xq, nq_ = conj(s) * xp + c * xq;
        if (j == p) { np_.im = T(0); }                                               // the diagonal stays real, the
        if (j == q) { nq_.im = T(0); np_ = cx<T>(T(0), T(0)); }                      // annihilated pair exactly zero
        if (j == p) nq_ = cx<T>(T(0), T(0));
        A[p * ld + j] = np_; A[q * ld + j] = nq_;
      }
      __syncthreads();
    }
  }
}

// mode EIGH:     in0 = A;            out0 = evals (batch, n) ascending, out1 = evecs (batch, n, n) columns (may be null)
// mode SQRTM:    in0 = A (PSD);      out0 = sqrt(A) (batch, n, n)     [negative round-off eigenvalues clamp to 0]
// mode FIDELITY: in0 = rho, in1 = sigma;  out0 = F (batch) real = (tr sqrt(sqrt(rho) sigma sqrt(rho)))^2
template <typename TIO, int MODE>
__global__ void __launch_bounds__(NT) spectral_kernel(const cx<TIO>* __restrict__ in0, const cx<TIO>* __restrict__ in1,
                                                     void* __restrict__ out0, cx<TIO>* __restrict__ out1, int n, int want_v) {
  typedef double T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ T rot_c[NMAX / 2];
  __shared__ cx<T> rot_s[NMAX / 2];
  __shared__ int rot_p[NMAX / 2], rot_q[NMAX / 2];
  __shared__ T red[NT / 32];
  __shared__ T lam[NMAX];
  __shared__ int flag;
  const int ld = n | 1;
  cx<T>* A = reinterpret_cast<cx<T>*>(smem_raw);
  cx<T>* V = A + (size_t)n * ld;
  cx<T>* S = V + (size_t)n * ld;                                                     // FIDELITY only
  const size_t b = blockIdx.x, nn = (size_t)n * n;
  const bool use_v = MODE != MODE_EIGH || want_v;
  for (int e = threadIdx.x; e < n * n; e += NT) {
    const int r = e / n, c = e - r * n;
    const cx<TIO> a = in0[b * nn + e];
    A[r * ld + c] = cx<T>((T)a.re, (T)a.im);
    if (use_v) V[r * ld + c] = cx<T>(r == c ? T(1) : T(0), T(0));
  }
  __syncthreads();
  jacobi<T>(A, use_v ? V : nullptr, n, ld, rot_c, rot_s, rot_p, rot_q, red, &flag);
  __syncthreads();
  if (MODE == MODE_EIGH) {
    TIO* evals = reinterpret_cast<TIO*>(out0) + b * n;
    for (int k = threadIdx.x; k < n; k += NT) lam[k] = A[k * ld + k].re;
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += NT) {                                      // rank sort: ascending, stable
      const T v = lam[k];
      int rank = 0;
      for (int j = 0; j < n; ++j) rank += (lam[j] < v || (lam[j] == v && j < k)) ? 1 : 0;
      if (!(v == v)) rank = k;                                                       // NaN: keep positions
      evals[rank] = (TIO)v;
    }
    if (out1 && want_v) {
      for (int e = threadIdx.x; e < n * n; e += NT) {
        const int i = e / n, k = e - i * n;
        const T v = lam[k];
        int rank = 0;
        for (int j = 0; j < n; ++j) rank += (lam[j] < v || (lam[j] == v && j < k)) ? 1 : 0;
        if (!(v == v)) rank = k;
        out1[b * nn + (size_t)i * n + rank] = cx<TIO>((TIO)V[i * ld + k].re, (TIO)V[i * ld + k].im);
      }
    }
    return;
  }
  // f(lambda) = sqrt(lambda) above the round-off floor, 0 below; NaN propagates
  {
    T mx = T(0);
    for (int k = 0; k < n; ++k) mx = fmax(mx, A[k * ld + k].re);
    const T floor_ = noise_floor<TIO>(n) * mx;
    for (int k = threadIdx.x; k < n; k += NT) { const T v = A[k * ld + k].re; lam[k] = v > floor_ ? sqrt(v) : (v == v ? T(0) : v); }
  }
  __syncthreads();
  cx<T>* R = MODE == MODE_SQRTM ? A : S;                                             // sqrt(in0) = V f(lambda) V^H
  for (int e = threadIdx.x; e < n * n; e += NT) {
    const int i = e / n, j = e - i * n;
    cx<T> acc(T(0), T(0));
    for (int k = 0; k < n; ++k) { const cx<T> w = lam[k] * V[i * ld + k]; cfma(acc, w, conj(V[j * ld + k])); }
    R[i * ld + j] = acc;
  }
  __syncthreads();
  if (MODE == MODE_SQRTM) {
    cx<TIO>* o = reinterpret_cast<cx<TIO>*>(out0) + b * nn;
    for (int e = threadIdx.x; e < n * n; e += NT) { const int i = e / n, j = e - i * n; o[e] = cx<TIO>((TIO)A[i * ld + j].re, (TIO)A[i * ld + j].im); }
    return;
  }
  // FIDELITY: M = S sigma S (Hermitian PSD) -> eigenvalues mu -> F = (sum sqrt(max(mu, 0)))^2
  for (int e = threadIdx.x; e < n * n; e += NT) {
    const int r = e / n, c = e - r * n;
    const cx<TIO> a = in1[b * nn + e];
    A[r * ld + c] = cx<T>((T)a.re, (T)a.im);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < n * n; e += NT) {                                    // V <- sigma S
    const int i = e / n, j = e - i * n;
    cx<T> acc(T(0), T(0));
    for (int k = 0; k < n; ++k) cfma(acc, A[i * ld + k], S[k * ld + j]);
    V[i * ld + j] = acc;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < n * n; e += NT) {                                    // A <- S V, symmetrised on the fly
    const int i = e / n, j = e - i * n;
    if (j < i) continue;
    cx<T> a(T(0), T(0)), bb(T(0), T(0));
    for (int k = 0; k < n; ++k) { cfma(a, S[i * ld + k], V[k * ld + j]); cfma(bb, S[j * ld + k], V[k * ld + i]); }
    const cx<T> h(T(0.5) * (a.re + bb.re), i == j ? T(0) : T(0.5) * (a.im - bb.im));
    A[i * ld + j] = h; A[j * ld + i] = conj(h);
  }
  __syncthreads();
  jacobi<T>(A, nullptr, n, ld, rot_c, rot_s, rot_p, rot_q, red, &flag);
  __syncthreads();
  T part = T(0), mx = T(0);
  for (int k = 0; k < n; ++k) mx = fmax(mx, A[k * ld + k].re);
  const T floor_ = noise_floor<TIO>(n) * mx;
  for (int k = threadIdx.x; k < n; k += NT) { const T v = A[k * ld + k].re; part += v > floor_ ? sqrt(v) : (v == v ? T(0) : v); }
  part = block_sum(part, red);
  if (threadIdx.x == 0) reinterpret_cast<TIO*>(out0)[b] = (TIO)(part * part);
}

// stats[b] = { number of entries with a_rc != conj(a_cr)  (exact test, as jnp.all(o == dag(o)) at qgrad_qutip.py:367),
//              Re tr A, Im tr A, max_rc |a_rc - conj(a_cr)| }            one CTA per matrix, any n
template <typename T>
__global__ void __launch_bounds__(NT) herm_stats_kernel(const cx<T>* __restrict__ Ain, T* __restrict__ stats, int n) {
  __shared__ T red[NT / 32];
  const size_t b = blockIdx.x, nn = (size_t)n * n;
  const cx<T>* A = Ain + b * nn;
  T bad = T(0), tr_re = T(0), tr_im = T(0), dmax = T(0);
  for (size_t e = threadIdx.x; e < nn; e += NT) {
    const int r = (int)(e / n), c = (int)(e % n);
    const cx<T> v = A[e], w = A[(size_t)c * n + r];
    if (!(v.re == w.re && v.im == -w.im)) bad += T(1);                               // NaN counts as a mismatch, as == does
    dmax = fmax(dmax, fmax(fabs(v.re - w.re), fabs(v.im + w.im)));
    if (r == c) { tr_re += v.re; tr_im += v.im; }
  }
  bad = block_sum(bad, red); tr_re = block_sum(tr_re, red); tr_im = block_sum(tr_im, red);
  dmax = warp_max(dmax);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dmax;
  __syncthreads();
  if (threadIdx.x == 0) {
    T m = T(0);
    for (int w = 0; w < NT / 32; ++w) m = fmax(m, red[w]);
    stats[b * 4 + 0] = bad; stats[b * 4 + 1] = tr_re; stats[b * 4 + 2] = tr_im; stats[b * 4 + 3] = m;
  }
}

template <typename T>
static int run(int mode, const void* in0, const void* in1, void* out0, void* out1, int64_t batch, int n, cudaStream_t st) {
  if (batch > 0x7fffffffLL) return QG_ERR_ARG;
  if (mode == 3) {
    herm_stats_kernel<T><<<(unsigned)batch, NT, 0, st>>>((const cx<T>*)in0, (T*)out0, n);
    QG_CHECK_LAUNCH();
    return QG_OK;
  }
  if (n > NMAX) return QG_ERR_DIM;
  const int ld = n | 1;
  const size_t smem = (size_t)(mode == MODE_FIDELITY ? 3 : 2) * n * ld * sizeof(cx<double>);
  static std::atomic<unsigned long long> attr_done{0};
  if (first_use_on_device(attr_done)) {
    const int cap = 3 * NMAX * (NMAX | 1) * (int)sizeof(cx<double>);
    QG_RETURN_IF_CUDA(cudaFuncSetAttribute(spectral_kernel<T, MODE_EIGH>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
    QG_RETURN_IF_CUDA(cudaFuncSetAttribute(spectral_kernel<T, MODE_SQRTM>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
    QG_RETURN_IF_CUDA(cudaFuncSetAttribute(spectral_kernel<T, MODE_FIDELITY>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
  }
  switch (mode) {
    case MODE_EIGH:
      spectral_kernel<T, MODE_EIGH><<<(unsigned)batch, NT, smem, st>>>((const cx<T>*)in0, nullptr, out0, (cx<T>*)out1, n, out1 ? 1 : 0);
      break;
    case MODE_SQRTM:
      spectral_kernel<T, MODE_SQRTM><<<(unsigned)batch, NT, smem, st>>>((const cx<T>*)in0, nullptr, out0, nullptr, n, 1);
      break;
    default:
      spectral_kernel<T, MODE_FIDELITY><<<(unsigned)batch, NT, smem, st>>>((const cx<T>*)in0, (const cx<T>*)in1, out0, nullptr, n, 1);
  }
  QG_CHECK_LAUNCH();
  return QG_OK;
}

}  // namespace spectral

int spectral_op(int dtype, int mode, const void* in0, const void* in1, void* out0, void* out1, int64_t batch, int n, cudaStream_t st) {
  return dtype == QG_C128 ? spectral::run<double>(mode, in0, in1, out0, out1, batch, n, st)
                          : spectral::run<float>(mode, in0, in1, out0, out1, batch, n, st);
}

}  // namespace qg

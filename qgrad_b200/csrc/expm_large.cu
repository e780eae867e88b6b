// K2 / K3 (large dimension, n > 32): batched complex expm and its Frechet pull-back as a chain
// of tiled complex GEMMs over a caller-provided workspace (see gemm.cuh for the tensor-pipe
// kernel).  Replaces scipy.linalg.expm (qgrad/qgrad_qutip.py:280,
// examples/Unitary-Learning-no-qgrad.py:137) at the north star's "large dimension" sizes.
//
// Per matrix, on the device: the top Pade order always ([7/7] for complex128, [5/5] for complex64),
// As = A/2^s zero-padded to np = ceil(n/64)*64 (exp(diag(A,0)) = diag(exp A, I)).
//
// Number of squarings s.  General matrices: from ||A||_1 against the order's bound (Higham 2005 /
// Al-Mohy & Higham 2009).  Matrices detected as Hermitian or skew-Hermitian (an O(n^2) test fused
// into the norm kernel; every Unitary(H)(t) = expm(-iHt) input is one) are NORMAL, so
// ||A||_2 = rho(A) <= ||A^k||_1^(1/k) for every k: the same truncation bounds hold in the 2-norm with
// d_k = ||A^k||_1^(1/k) (k = 6, or 4 for complex64) in place of ||A||_1, the powers being the ones
// the Pade evaluation needs anyway (computed before the final scaling, then rescaled by
// 2^-2e, 2^-4e, 2^-6e).  For a GUE-like H of dimension 1024 this is ||A||_2 ~ 2.6 against
// ||A||_1 ~ 30: three squarings (nine n^3 products of the fused pass) fewer, and a smaller
// squaring error.  ||q(A) - I||_2 <= e^(0.4) - 1 < 1 keeps the field of values of the Pade denominator
// in the right half plane, so the unpivoted elimination stays stable.
// d_k over-estimates rho by up to n^(1/2k); a power iteration on the top even power (power_iter_kernel) tightens
// it where that can save a squaring level (the learning step's Pauli Hamiltonians: s = 5 -> 4).
//   forward : A2, A4, A6, [d_6 and the rho estimate -> s, rescale, W], U-GEMM -> Q = V-U, P = V+U, Qi = Q^-1 by a
//             blocked (64) Gauss-Jordan sweep built from the same GEMM kernel (one or two complex128 matrices:
//             series of 1/q_7 + one Newton-Schulz step, series_inverse), R0 = Qi P,
//             R_{k+1} = R_k^2 (launched max_sq times, predicated per matrix on k < s[b]).
//   backward: E = Ybar^H/2^s, M2, M4, M6 (+Lw), Lu-GEMM -> D1 = Lu+Lv, D2 = Lu-Lv,
//             G = D1 + D2 R0, L0 = Qi G, L_{k+1} = R_k L_k + L_k R_k, Abar = L_s^H.
//   backward_skew (skew-Hermitian A, QG_ADJ_SKEW_BODY): the skew-Hermitian part of Abar from the body-frame
//             cotangent Z = Ybar U^H; every dual product has a Hermitian or skew-Hermitian result.
// Forward products 5 + s, backward 10 + 2 s (+1 for the inverse): 16 + 3 s GEMM-equivalents; with the structure
// hints 4t + 2 + s forward and 8t + 1 + t + s(1 + t) for backward_skew, t = (T+1)/2T of a product (T = np/64).
#include "gemm.cuh"

namespace qg {
namespace large {

constexpr int NBK = 64;                       // Gauss-Jordan block = GEMM tile of the FMA / DMMA kernel
// width the per-matrix scratch (block inverse, saved pivot column) is sized for: complex64 may run its sweep in 128-wide
// blocks on the tensor-core kernel (block_inverse_tc)
template <typename T> __host__ __device__ constexpr int scratch_bk() { return sizeof(T) == 4 ? 128 : 64; }

// process-wide planner switch (qg_set_plan_rigorous / QG_PLAN_RIGOROUS): 1 = no spectral-radius ESTIMATE, the number of
// squarings of normal matrices comes from rigorous bounds only (||A||_1 and d_k = ||A^k||_1^(1/k))
static std::atomic<int> g_plan_rigorous{-1};
static bool plan_rigorous() {
  int v = g_plan_rigorous.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("QG_PLAN_RIGOROUS");
    v = (e && *e && *e != '0') ? 1 : 0;
    g_plan_rigorous.store(v, std::memory_order_relaxed);
  }
  return v != 0;
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
__host__ __device__ inline size_t align_up_dev(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int pad_dim(int n) { return (n + 63) / 64 * 64; }

// forward buffers
enum { B_AS = 0, B_A2, B_A4, B_A6, B_W, B_Q, B_P, B_FWD_COUNT };
// backward buffers (after the chain)
enum { D_E = 0, D_M2, D_M4, D_M6, D_LW, D_COUNT };

template <typename T>
struct Layout {
  int n, np, max_sq, chain;      // chain = number of R buffers
  int64_t batch;
  size_t mat_elems;              // np*np
  int* s; int* flag; int* sext; int* normal;   // squarings, overflow flag, late extra scaling, normal-matrix flag
  int* sym;                                    // symmetry class: 0 general, 1 skew-Hermitian, 2 Hermitian
  T* asym; T* n1;                              // (skew dist, herm dist, sum |a|) and ||A||_1 per matrix
  T* colsum; cx<T>* dinv; cx<T>* colsave; cx<T>* big;   // colsave: np x 64 per matrix (old pivot column of the inverse)
  __host__ __device__ cx<T>* buf(int i) const { return big + (size_t)i * batch * mat_elems; }
  __host__ __device__ cx<T>* fwd(int i) const { return buf(i); }
  __host__ __device__ cx<T>* R(int k) const { return buf(B_FWD_COUNT + (k % chain)); }
  __host__ __device__ cx<T>* dual(int i) const { return buf(B_FWD_COUNT + chain + i); }
};

template <typename T>
static size_t per_matrix_bytes(int n, int mode, int max_sq) {
  const int np = pad_dim(n);
  const int chain = mode == QG_EXPM_FWD_VJP ? max_sq + 1 : 2;
  const int nbuf = B_FWD_COUNT + chain + (mode == QG_EXPM_FWD_VJP ? D_COUNT : 0);
  return 256 /* s, flag */ + align_up((size_t)np * sizeof(T), 256) + align_up((size_t)scratch_bk<T>() * scratch_bk<T>() * sizeof(cx<T>), 256) +
         align_up((size_t)np * scratch_bk<T>() * sizeof(cx<T>), 256) + (size_t)nbuf * align_up((size_t)np * np * sizeof(cx<T>), 256);
}

template <typename T>
static size_t workspace_bytes(int n, int64_t batch, int mode, int max_sq) {
  return 256 + (size_t)batch * per_matrix_bytes<T>(n, mode, max_sq);
}

template <typename T>
static bool carve(Layout<T>& L, void* ws, size_t ws_bytes, int n, int64_t batch, int mode, int max_sq) {
  if (!ws || ws_bytes < workspace_bytes<T>(n, batch, mode, max_sq)) return false;
  L.n = n; L.np = pad_dim(n); L.max_sq = max_sq; L.batch = batch;
  L.chain = mode == QG_EXPM_FWD_VJP ? max_sq + 1 : 2;
  L.mat_elems = (size_t)L.np * L.np;
  char* p = (char*)align_up((size_t)ws, 256);
  L.s = (int*)p; L.flag = L.s + batch; L.sext = L.flag + batch; L.normal = L.sext + batch; L.sym = L.normal + batch;
  L.asym = (T*)(p + (size_t)batch * 32); L.n1 = L.asym + 3 * batch;      // 32 + 4*sizeof(T) <= 64 of the 256 B per matrix
  p += (size_t)batch * 256;
  L.colsum = (T*)p; p += (size_t)batch * align_up((size_t)L.np * sizeof(T), 256);
  L.dinv = (cx<T>*)p; p += (size_t)batch * align_up((size_t)scratch_bk<T>() * scratch_bk<T>() * sizeof(cx<T>), 256);
  L.colsave = (cx<T>*)p; p += (size_t)batch * align_up((size_t)L.np * scratch_bk<T>() * sizeof(cx<T>), 256);
  L.big = (cx<T>*)p;
  return true;
}

// ------------------------------------------------------------------ small helper kernels
// partial column abs-sums of X (n x n, row pitch ld): grid (ceil(n/64), ceil(n/64), batch), block 256 =
// 64 cols x 4 row lanes.  With asym != null it also accumulates, per matrix, the entrywise distances
// sum |x_rc + conj(x_cr)| (to skew-Hermitian), sum |x_rc - conj(x_cr)| (to Hermitian) and sum |x_rc|.
template <typename T>
__global__ void colsum_kernel(const cx<T>* __restrict__ X, T* __restrict__ colsum, T* __restrict__ asym, int n, int ld,
                              size_t mat_stride, int np) {
  __shared__ T part[4][64];
  __shared__ T red3[3][8];
  const int c = blockIdx.x * 64 + (threadIdx.x & 63), rl = threadIdx.x >> 6;
  const int r0 = blockIdx.y * 64;
  const cx<T>* a = X + (size_t)blockIdx.z * mat_stride;
  T s = T(0), d_skew = T(0), d_herm = T(0);
  if (c < n)
    for (int r = r0 + rl; r < min(r0 + 64, n); r += 4) {
      const cx<T> v = a[(size_t)r * ld + c];
      s += cabs_(v);
      if (asym) {
        const cx<T> w = a[(size_t)c * ld + r];
        d_skew += cabs_(cx<T>(v.re + w.re, v.im - w.im));
        d_herm += cabs_(cx<T>(v.re - w.re, v.im + w.im));
      }
    }
  part[rl][threadIdx.x & 63] = s;
  if (asym) {
    T t0 = warp_sum(d_skew), t1 = warp_sum(d_herm), t2 = warp_sum(s);
    if ((threadIdx.x & 31) == 0) { red3[0][threadIdx.x >> 5] = t0; red3[1][threadIdx.x >> 5] = t1; red3[2][threadIdx.x >> 5] = t2; }
  }
  __syncthreads();
  if (threadIdx.x < 64 && c < n) {
    const T tot = part[0][threadIdx.x] + part[1][threadIdx.x] + part[2][threadIdx.x] + part[3][threadIdx.x];
    atomicAdd(&colsum[(size_t)blockIdx.z * (((size_t)np * sizeof(T) + 255) / 256 * 256 / sizeof(T)) + c], tot);
  }
  if (asym && threadIdx.x < 3) {
    T tot = T(0);
    for (int w = 0; w < 8; ++w) tot += red3[threadIdx.x][w];
    atomicAdd(&asym[(size_t)blockIdx.z * 3 + threadIdx.x], tot);
  }
}

// how many squarings a normal matrix can save at most against the 1-norm rule: ||A||_1 <= sqrt(n) ||A||_2
constexpr int kNormalGain = 6;

// max over the n column sums of one matrix (NaN -> inf), valid in thread 0
template <typename T>
__device__ __forceinline__ T block_colmax(const T* cs, int n, T* red) {
  T best = T(0);
  bool isnan_ = false;
  for (int c = threadIdx.x; c < n; c += blockDim.x) { const T v = cs[c]; if (v != v) isnan_ = true; best = fmax(best, v); }
  if (isnan_) best = (T)INFINITY;
  best = warp_max(best);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
  __syncthreads();
  T m = T(0);
  if (threadIdx.x == 0)
    for (int w = 0; w < (int)(blockDim.x + 31) / 32; ++w) m = fmax(m, red[w]);
  return m;
}

// Stage 1: ||A||_1, the 1-norm rule's s, and the normal-matrix test.  A normal matrix that needs any
// scaling is pre-scaled by kNormalGain levels less; stage 2 settles its s from the power norm.
template <typename T>
__global__ void plan_kernel(Layout<T> L, int frechet) {
  __shared__ T red[32];
  const int b = blockIdx.x;
  const size_t pitch = ((size_t)L.np * sizeof(T) + 255) / 256 * 256 / sizeof(T);
  const T m = block_colmax(L.colsum + (size_t)b * pitch, L.n, red);
  if (threadIdx.x == 0) {
    int mm, ss;
    pade_plan<T>((double)m, frechet != 0, &mm, &ss);      // the large path always evaluates the top order
    // Two tests.  NORMAL (for the choice of s only: the 2-norm rule tolerates a perturbation of this size): entrywise
    // distance to (skew-)Hermitian <= 2^-40 / 2^-18 of sum |a|.  STRUCTURED (sym: upper-triangle products mirrored into
    // the lower triangle -- which DROPS any non-Hermitian part): distance <= 4 u sum |a|, i.e. Hermitian up to the
    // rounding of building H from a formula; H - i gamma / 2 with a small gamma keeps its full products.
    const T tol = sizeof(T) == 8 ? T(9.1e-13) : T(3.8e-6);  // 2^-40, 2^-18
    const T tol_exact = sizeof(T) == 8 ? T(8.9e-16) : T(4.8e-7);  // 4 u
    const T* as = L.asym + (size_t)b * 3;
    const int near = (m == m && as[0] <= tol * as[2]) ? 1 : ((m == m && as[1] <= tol * as[2]) ? 2 : 0);
    const int sym = (near == 1 && as[0] <= tol_exact * as[2]) ? 1 : ((near == 2 && as[1] <= tol_exact * as[2]) ? 2 : 0);
    const bool normal = ss > 0 && ss < 64 && near != 0;
    L.sym[b] = sym;
    L.n1[b] = m;
    L.normal[b] = normal ? 1 : 0;
    L.sext[b] = 0;
    int bad = 0;
    if (normal) ss = ss > kNormalGain ? ss - kNormalGain : 0;
    else if (ss > L.max_sq) { bad = 1; ss = L.max_sq; }
    L.s[b] = ss;
    L.flag[b] = bad;
  }
}

// Stage 2a (normal matrices): power iteration on the top even power B = (2^-s A)^k, which is Hermitian up to sign,
// so ||B x|| / ||x|| <= rho(B) = rho^k and converges to it from below.  d_k over-estimates rho by up to
// n^(1/2k) (measured 1.2-1.25 on the 10-qubit Pauli Hamiltonians of the learning step: d_6 = 13.7 for rho = 11.0,
// which costs a whole squaring level); kPowerIters products with B -- degree 2k * kPowerIters in the spectrum --
// come within 2.5 % of rho on every spectrum tried (Pauli sums, GUE, clustered, isolated top), and the planner
// inflates the estimate by kPowerMargin.  SciPy's expm sizes its scaling from norm *estimates* as well
// (onenormest); an estimate 1 % short would raise the Pade truncation error by 1.01^15.
// One launch per iteration, grid (np / 8, batch), a warp per row; x lives in the (idle) colsave area, two vectors
// per matrix, stored un-normalised: every CTA recomputes ||x|| itself, in a fixed order (no atomics -> the
// planned s is reproducible).  Iteration 0 starts from a fixed pseudo-random complex vector.
// TWO start vectors run side by side (one pass over B serves both, the iteration is memory bound): a power iteration
// cannot see an eigenvector its start vector is orthogonal to, and with a single fixed vector such a matrix is easy to
// write down; the estimate is the larger of the two.  Fooling both takes a dominant eigenvector orthogonal to two fixed
// pseudo-random vectors -- the planner then falls back on the rigorous floor d_k n^(-1/2k) <= rho (the [7/7] truncation
// error grows by at most n^(15/12), tests/test_gpu_parity.py::test_planner_adversarial_spectra), and
// qg_set_plan_rigorous(1) / QG_PLAN_RIGOROUS=1 switches the estimate off altogether (s from d_k, a rigorous bound).
constexpr int kPowerIters = 8;
constexpr double kPowerMargin = 1.03;
__device__ __forceinline__ float hash_unit(unsigned v) {
  v ^= v >> 16; v *= 0x7feb352du; v ^= v >> 15; v *= 0x846ca68bu; v ^= v >> 16;
  return (float)(v & 0xffffu) * (1.0f / 32768.0f) - 1.0f;
}
template <typename T> __device__ __forceinline__ cx<T> start_vec(int which, int c) {
  const unsigned o = which ? 0x9e3779b9u : 0u;
  return cx<T>((T)hash_unit(2u * (unsigned)c + 1u + o), (T)hash_unit(2u * (unsigned)c + 2u + o));
}
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  T tot = T(0);
  for (int w = 0; w < (int)(blockDim.x + 31) / 32; ++w) tot += red[w];
  return tot;
}
// Gate: the iteration can only pay where some rho in [d_k n^(-1/2k), d_k] needs fewer levels than d_k itself
// (about half of all matrices at n = 64, four in five at n = 1024); the others keep normal[b] = 1 and are skipped.
template <typename T>
__global__ void power_gate_kernel(Layout<T> L, int frechet, int k) {
  __shared__ T red[32];
  const int b = blockIdx.x;
  if (!L.normal[b]) return;
  const size_t pitch = ((size_t)L.np * sizeof(T) + 255) / 256 * 256 / sizeof(T);
  const T m = block_colmax(L.colsum + (size_t)b * pitch, L.np, red);
  if (threadIdx.x == 0 && m == m && !isinf((double)m)) {
    const double bound = PadePlan<T>::bound(PadePlan<T>::kTop, frechet != 0);
    const double dk = pow((double)m, 1.0 / k);
    const double n1s = ldexp((double)L.n1[b], -L.s[b]);
    const double hi = dk < n1s ? dk : n1s, lo = dk * pow((double)L.np, -0.5 / k);
    int eh = 0, el = 0;
    for (double x = hi; !(x <= bound) && eh < 64; x *= 0.5) ++eh;
    for (double x = lo; !(x <= bound) && el < 64; x *= 0.5) ++el;
    if (el < eh) { L.normal[b] = 2; L.sext[b] = eh; }               // sext: scratch until plan_power_kernel sets it
  }
}
// ||B x|| / ||x|| is a lower bound on rho^k at every iteration: once it needs as many levels as d_k does, nothing can
// be gained and the matrix drops out (normal[b] back to 1).  Every CTA of a matrix takes the same decision from the
// same numbers; one thread records it for the launches that follow.
template <typename T>
__device__ __forceinline__ bool power_hopeless(const Layout<T>& L, int b, T nn, int frechet, int k) {
  if (!(nn > T(0)) || nn != nn || isinf((double)nn)) return false;
  const double bound = PadePlan<T>::bound(PadePlan<T>::kTop, frechet != 0);
  int el = 0;
  for (double x = pow((double)nn, 0.5 / k); !(x <= bound) && el < 64; x *= 0.5) ++el;
  return el >= L.sext[b];
}

// WPR warps share a row (8 / WPR rows per CTA): with one or two matrices in the batch -- the 8- and 4-rank shards of the
// learning step -- a warp per row leaves 7 warps per SM to stream a 16 MB matrix (23 us per product, 8 products per step);
// four warps per row give the memory system four times the loads in flight.
template <typename T, int WPR>
__global__ void __launch_bounds__(256) power_iter_kernel(Layout<T> L, const cx<T>* __restrict__ Bm, int it, int frechet, int k) {
  __shared__ T red[32];
  __shared__ T part[8][4];
  const int b = blockIdx.y;
  if (L.normal[b] != 2) return;
  const int np = L.np;
  // four vectors per matrix in the (idle) colsave area: [parity][which][np]
  cx<T>* vecs = L.colsave + (size_t)b * (align_up_dev((size_t)np * scratch_bk<T>() * sizeof(cx<T>), 256) / sizeof(cx<T>));
  const cx<T>* xin = vecs + (size_t)(it & 1) * 2 * np;
  cx<T>* xout = vecs + (size_t)((it + 1) & 1) * 2 * np;
  auto x_at = [&](int which, int c) { return it == 0 ? start_vec<T>(which, c) : xin[(size_t)which * np + c]; };
  T n0 = T(0), n1 = T(0);
  for (int c = threadIdx.x; c < np; c += blockDim.x) {
    const cx<T> v = x_at(0, c), w = x_at(1, c);
    n0 += v.re * v.re + v.im * v.im; n1 += w.re * w.re + w.im * w.im;
  }
  n0 = block_sum(n0, red); n1 = block_sum(n1, red);
  if (it > 0 && power_hopeless(L, b, fmax(n0, n1), frechet, k)) {
    if (blockIdx.x == 0 && threadIdx.x == 0) L.normal[b] = 1;
    return;
  }
  const T i0 = n0 > T(0) && n0 == n0 && n0 < (T)INFINITY ? T(1) / sqrt(n0) : T(0);
  const T i1 = n1 > T(0) && n1 == n1 && n1 < (T)INFINITY ? T(1) / sqrt(n1) : T(0);
  constexpr int RPC = 8 / WPR;                                        // rows per CTA
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rl = warp / WPR, seg = warp % WPR;
  const int r = blockIdx.x * RPC + rl;
  const cx<T>* row = Bm + (size_t)b * L.mat_elems + (size_t)r * np;
  cx<T> a0(T(0), T(0)), a1(T(0), T(0));
  const int c_lo = seg * (np / WPR), c_hi = c_lo + np / WPR;          // np is a multiple of 64
  for (int c = c_lo + lane; c < c_hi; c += 32) { const cx<T> m = row[c]; cfma(a0, m, x_at(0, c)); cfma(a1, m, x_at(1, c)); }
  a0.re = warp_sum(a0.re); a0.im = warp_sum(a0.im); a1.re = warp_sum(a1.re); a1.im = warp_sum(a1.im);
  if (WPR == 1) {
    if (lane == 0) { xout[r] = cx<T>(a0.re * i0, a0.im * i0); xout[(size_t)np + r] = cx<T>(a1.re * i1, a1.im * i1); }
  } else {
    if (lane == 0) { part[warp][0] = a0.re; part[warp][1] = a0.im; part[warp][2] = a1.re; part[warp][3] = a1.im; }
    __syncthreads();
    if (seg == 0 && lane == 0) {
      T s0 = T(0), s1 = T(0), s2 = T(0), s3 = T(0);
#pragma unroll
      for (int q = 0; q < WPR; ++q) { s0 += part[warp + q][0]; s1 += part[warp + q][1]; s2 += part[warp + q][2]; s3 += part[warp + q][3]; }
      xout[r] = cx<T>(s0 * i0, s1 * i0); xout[(size_t)np + r] = cx<T>(s2 * i1, s3 * i1);
    }
  }
}

// The same iteration with ONE CTA per matrix running all `iters` products (x ping-pongs in shared memory, B is
// re-read through L1/L2): for batches that fill the SMs on their own.  At n = 64, batch 2288 the per-iteration
// launches cost 8 x 49 us (7.5 % of expm+VJP); this is one launch.  Writes x_iters where plan_power_kernel reads it.
template <typename T>
__global__ void __launch_bounds__(256, 4) power_iter_fused_kernel(Layout<T> L, const cx<T>* __restrict__ Bm, int iters, int frechet, int k) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ T red[32];
  __shared__ T red2[2][8];
  const int b = blockIdx.x;
  if (L.normal[b] != 2) return;
  const int np = L.np;
  cx<T>* xs = reinterpret_cast<cx<T>*>(smem_raw);                    // [parity][which][np]
  const cx<T>* Bb = Bm + (size_t)b * L.mat_elems;
  T n0 = T(0), n1 = T(0);
  for (int c = threadIdx.x; c < np; c += blockDim.x) {
    const cx<T> v = start_vec<T>(0, c), w = start_vec<T>(1, c);
    xs[c] = v; xs[np + c] = w;
    n0 += v.re * v.re + v.im * v.im; n1 += w.re * w.re + w.im * w.im;
  }
  n0 = block_sum(n0, red); n1 = block_sum(n1, red);                  // their barriers also publish xs
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int it = 0; it < iters; ++it) {
    const cx<T>* xin = xs + (size_t)(it & 1) * 2 * np;
    cx<T>* xout = xs + (size_t)((it + 1) & 1) * 2 * np;
    if (it > 0 && power_hopeless(L, b, fmax(n0, n1), frechet, k)) {
      if (threadIdx.x == 0) L.normal[b] = 1;
      return;
    }
    const T i0 = n0 > T(0) && n0 == n0 && n0 < (T)INFINITY ? T(1) / sqrt(n0) : T(0);
    const T i1 = n1 > T(0) && n1 == n1 && n1 < (T)INFINITY ? T(1) / sqrt(n1) : T(0);
    T m0 = T(0), m1 = T(0);                                          // sums |y_r|^2 over this warp's rows (lane 0)
    for (int r0 = warp * 2; r0 < np; r0 += nw * 2) {                 // two rows x two vectors: four independent chains
      const cx<T>* row = Bb + (size_t)r0 * np;
      cx<T> a00(T(0), T(0)), a01 = a00, a10 = a00, a11 = a00;
      for (int c = lane; c < np; c += 32) {
        const cx<T> x0 = xin[c], x1 = xin[np + c], ra = row[c], rb = row[np + c];
        cfma(a00, ra, x0); cfma(a01, ra, x1); cfma(a10, rb, x0); cfma(a11, rb, x1);
      }
      a00.re = warp_sum(a00.re); a00.im = warp_sum(a00.im); a01.re = warp_sum(a01.re); a01.im = warp_sum(a01.im);
      a10.re = warp_sum(a10.re); a10.im = warp_sum(a10.im); a11.re = warp_sum(a11.re); a11.im = warp_sum(a11.im);
      if (lane == 0) {
        const cx<T> y00(a00.re * i0, a00.im * i0), y10(a10.re * i0, a10.im * i0), y01(a01.re * i1, a01.im * i1), y11(a11.re * i1, a11.im * i1);
        xout[r0] = y00; xout[r0 + 1] = y10; xout[np + r0] = y01; xout[np + r0 + 1] = y11;
        m0 += y00.re * y00.re + y00.im * y00.im; m0 += y10.re * y10.re + y10.im * y10.im;
        m1 += y01.re * y01.re + y01.im * y01.im; m1 += y11.re * y11.re + y11.im * y11.im;
      }
    }
    __syncthreads();                                                 // everyone is done with red2's previous values
    if (lane == 0) { red2[0][warp] = m0; red2[1][warp] = m1; }
    __syncthreads();
    n0 = n1 = T(0);
    for (int w = 0; w < nw; ++w) { n0 += red2[0][w]; n1 += red2[1][w]; }   // fixed order: reproducible
  }
  cx<T>* vec = L.colsave + (size_t)b * (align_up_dev((size_t)np * scratch_bk<T>() * sizeof(cx<T>), 256) / sizeof(cx<T>)) + (size_t)(iters & 1) * 2 * np;
  const cx<T>* xfin = xs + (size_t)(iters & 1) * 2 * np;
  for (int c = threadIdx.x; c < 2 * np; c += blockDim.x) vec[c] = xfin[c];
}

// Stage 2b (normal matrices): d_k = ||(2^-s A)^k||_1^(1/k) >= rho = ||2^-s A||_2 and the power-iteration estimate
// of rho; e extra levels so that min(||2^-s A||_1, d_k, max(margin * estimate, d_k n^(-1/2k))) / 2^e <= bound
// (d_k n^(-1/2k) <= rho always: ||B||_1 <= sqrt(n) ||B||_2).
template <typename T>
__global__ void plan_power_kernel(Layout<T> L, int frechet, int k, int iters) {
  __shared__ T red[32];
  const int b = blockIdx.x;
  if (!L.normal[b]) return;
  const size_t pitch = ((size_t)L.np * sizeof(T) + 255) / 256 * 256 / sizeof(T);
  const T m = block_colmax(L.colsum + (size_t)b * pitch, L.np, red);
  T nn = T(0);
  if (iters > 0 && L.normal[b] == 2) {
    const cx<T>* x = L.colsave + (size_t)b * (align_up_dev((size_t)L.np * scratch_bk<T>() * sizeof(cx<T>), 256) / sizeof(cx<T>)) +
                     (size_t)(iters & 1) * 2 * L.np;
    T n0 = T(0), n1 = T(0);
    for (int c = threadIdx.x; c < L.np; c += blockDim.x) {
      const cx<T> v = x[c], w = x[L.np + c];
      n0 += v.re * v.re + v.im * v.im; n1 += w.re * w.re + w.im * w.im;
    }
    n0 = block_sum(n0, red); n1 = block_sum(n1, red);
    nn = fmax(n0, n1);                                               // both are lower bounds on rho^2k: keep the better one
  }
  if (threadIdx.x == 0) {
    const int s0 = L.s[b];
    const double dk = pow((double)m, 1.0 / k);
    const double n1s = ldexp((double)L.n1[b], -s0);
    double x = (dk < n1s) ? dk : n1s;                          // NaN compares false -> n1s
    if (m != m || isinf((double)m)) x = n1s;
    else if (iters > 0 && L.normal[b] == 2 && nn == nn && nn > T(0) && !isinf((double)nn)) {
      const double est = kPowerMargin * pow(sqrt((double)nn), 1.0 / k);
      const double floor_ = dk * pow((double)L.np, -0.5 / k);
      const double r2 = est > floor_ ? est : floor_;
      if (r2 < x) x = r2;
    }
    const double bound = PadePlan<T>::bound(PadePlan<T>::kTop, frechet != 0);
    int e = 0;
    while (!(x <= bound) && e < 64) { x *= 0.5; ++e; }
    int ss = s0 + e, bad = 0;
    if (ss > L.max_sq) { bad = 1; e = 0; ss = L.max_sq; }
    L.s[b] = ss; L.sext[b] = e; L.flag[b] = bad;
  }
}

// After stage 2: As *= 2^-e, A2 *= 2^-2e, A4 *= 2^-4e, A6 *= 2^-6e (e = 0: untouched) and
// W = c7 A6 + c5 A4 + c3 A2 + c1 I  (order 5: W = c5 A4 + c3 A2 + c1 I).
template <typename T>
__global__ void rescale_w_kernel(Layout<T> L, int m, T c1, T c3, T c5, T c7) {
  const int b = blockIdx.y;
  const int e = L.sext[b];
  const T f1 = (T)ldexp(1.0, -e), f2 = (T)ldexp(1.0, -2 * e), f4 = (T)ldexp(1.0, -4 * e), f6 = (T)ldexp(1.0, -6 * e);
  const size_t off = (size_t)b * L.mat_elems;
  cx<T>* As = L.fwd(B_AS) + off; cx<T>* A2 = L.fwd(B_A2) + off; cx<T>* A4 = L.fwd(B_A4) + off;
  cx<T>* A6 = L.fwd(B_A6) + off; cx<T>* W = L.fwd(B_W) + off;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < L.mat_elems; i += (size_t)gridDim.x * blockDim.x) {
    cx<T> a2 = A2[i], a4 = A4[i];
    if (e) {
      cx<T> a = As[i];
      As[i] = cx<T>(a.re * f1, a.im * f1);
      a2 = cx<T>(a2.re * f2, a2.im * f2); A2[i] = a2;
      a4 = cx<T>(a4.re * f4, a4.im * f4); A4[i] = a4;
    }
    const int r = (int)(i / L.np), c = (int)(i % L.np);
    cx<T> w(fma(c3, a2.re, r == c ? c1 : T(0)), c3 * a2.im);
    w.re = fma(c5, a4.re, w.re); w.im = fma(c5, a4.im, w.im);
    if (m == 7) {
      cx<T> a6 = A6[i];
      if (e) { a6 = cx<T>(a6.re * f6, a6.im * f6); A6[i] = a6; }
      w.re = fma(c7, a6.re, w.re); w.im = fma(c7, a6.im, w.im);
    }
    W[i] = w;
  }
}

// As = 2^-s A, zero padded to np x np.   grid (np/64, np/4?, batch) -> use 1-D over elements
template <typename T>
__global__ void scale_pad_kernel(const cx<T>* __restrict__ A, cx<T>* __restrict__ As, const int* __restrict__ s,
                                 int n, int np) {
  const int b = blockIdx.y;
  const T sc = (T)ldexp(1.0, -s[b]);
  const size_t total = (size_t)np * np;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / np), c = (int)(e % np);
    cx<T> v(T(0), T(0));
    if (r < n && c < n) { v = A[(size_t)b * n * n + (size_t)r * n + c]; v.re *= sc; v.im *= sc; }
    As[(size_t)b * total + e] = v;
  }
}

// E = 2^-s Ybar^H (conj_adj) or 2^-s Ybar^T, zero padded.  32x32 tile transpose through smem.
// grid (np/32, np/32, batch), block (32, 8)
template <typename T>
__global__ void scale_pad_adjoint_kernel(const cx<T>* __restrict__ Yb, cx<T>* __restrict__ E, const int* __restrict__ s,
                                         int n, int np, int conj_adj, int use_scale) {
  __shared__ cx<T> tile[32][33];
  const int b = blockIdx.z;
  const T sc = use_scale ? (T)ldexp(1.0, -s[b]) : T(1);
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;      // output tile origin (r0, c0)
  // read input tile at (c0.., r0..): input row = c0 + ty.., input col = r0 + tx
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int ir = c0 + i, ic = r0 + threadIdx.x;
    cx<T> v(T(0), T(0));
    if (ir < n && ic < n) v = Yb[(size_t)b * n * n + (size_t)ir * n + ic];
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int orow = r0 + i, ocol = c0 + threadIdx.x;
    cx<T> v = tile[threadIdx.x][i];
    if (conj_adj) v.im = -v.im;
    v.re *= sc; v.im *= sc;
    E[(size_t)b * np * np + (size_t)orow * np + ocol] = v;
  }
}

// E = 2^-s (Z - Z^H) / 2, zero padded: the skew-Hermitian part of the body-frame cotangent (backward_skew).
// grid (np/32, np/32, batch), block (32, 8)
template <typename T>
__global__ void skew_pad_kernel(const cx<T>* __restrict__ Z, cx<T>* __restrict__ E, const int* __restrict__ s, int n, int np) {
  __shared__ cx<T> tile[32][33];
  const int b = blockIdx.z;
  const T sc = (T)ldexp(0.5, -s[b]);
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const cx<T>* z = Z + (size_t)b * n * n;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int ir = c0 + i, ic = r0 + threadIdx.x;
    tile[i][threadIdx.x] = (ir < n && ic < n) ? z[(size_t)ir * n + ic] : cx<T>(T(0), T(0));
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int orow = r0 + i, ocol = c0 + threadIdx.x;
    const cx<T> w = tile[threadIdx.x][i];                                  // Z[ocol][orow]
    const cx<T> v = (orow < n && ocol < n) ? z[(size_t)orow * n + ocol] : cx<T>(T(0), T(0));
    E[(size_t)b * np * np + (size_t)orow * np + ocol] = cx<T>(sc * (v.re - w.re), sc * (v.im + w.im));
  }
}

// Abar[b] = S_final[b] cropped (no transpose); S_final lives in dual(D_E) for even s, dual(D_M2) for odd s.
// NaN for matrices that needed more squarings than allowed or that are not skew-Hermitian.
template <typename T>
__global__ void finalize_skew_kernel(Layout<T> L, cx<T>* __restrict__ Abar) {
  const int b = blockIdx.y;
  const cx<T>* src = ((L.s[b] & 1) ? L.dual(D_M2) : L.dual(D_E)) + (size_t)b * L.mat_elems;
  const T nanv = (L.flag[b] || L.sym[b] != 1) ? (T)nan("") : T(0);
  const size_t total = (size_t)L.n * L.n;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / L.n), c = (int)(e % L.n);
    cx<T> v = src[(size_t)r * L.np + c];
    v.re += nanv; v.im += nanv;
    Abar[(size_t)b * total + e] = v;
  }
}

// Y[b] = R_{s[b]}[b] cropped to n x n (NaN if the matrix needed more squarings than allowed)
template <typename T>
__global__ void finalize_kernel(Layout<T> L, cx<T>* __restrict__ Y) {
  const int b = blockIdx.y;
  const cx<T>* src = L.R(L.s[b]) + (size_t)b * L.mat_elems;
  const T nanv = L.flag[b] ? (T)nan("") : T(0);
  const size_t total = (size_t)L.n * L.n;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / L.n), c = (int)(e % L.n);
    cx<T> v = src[(size_t)r * L.np + c];
    v.re += nanv; v.im += nanv;
    Y[(size_t)b * total + e] = v;
  }
}

// Abar[b] = (L_final[b])^H or ^T cropped; L_final lives in dual(D_LW) for even s, dual(D_E) for odd s.
// grid (ceil(n/32), ceil(n/32), batch), block (32, 8)
template <typename T>
__global__ void finalize_adjoint_kernel(Layout<T> L, cx<T>* __restrict__ Abar, int conj_adj) {
  __shared__ cx<T> tile[32][33];
  const int b = blockIdx.z;
  const cx<T>* src = ((L.s[b] & 1) ? L.dual(D_E) : L.dual(D_LW)) + (size_t)b * L.mat_elems;
  const T nanv = L.flag[b] ? (T)nan("") : T(0);
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int ir = c0 + i, ic = r0 + threadIdx.x;           // always inside np x np
    tile[i][threadIdx.x] = src[(size_t)ir * L.np + ic];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int orow = r0 + i, ocol = c0 + threadIdx.x;
    if (orow < L.n && ocol < L.n) {
      cx<T> v = tile[threadIdx.x][i];
      if (conj_adj) v.im = -v.im;
      v.re += nanv; v.im += nanv;
      Abar[(size_t)b * L.n * L.n + (size_t)orow * L.n + ocol] = v;
    }
  }
}

// ------------------------------------------------------------------ blocked Gauss-Jordan pieces
// Dinv[b] = inverse of the 64x64 diagonal block kb of Q[b], also written back over the block: register-
// resident unpivoted Gauss-Jordan sweep (common.cuh gj_inverse_smem), 256 workers x (4 x 4) elements plus
// one extra warp whose lane 0 does nothing but the next pivot's reciprocal.
constexpr int kDiagWork = 256, kDiagThreads = kDiagWork + 32;
// MINB = 2 caps the registers at 113 so that two CTAs share an SM: a little slower per CTA (the complex128 build wants
// 118), nearly twice the throughput once there are more matrices than SMs.
template <typename T, int MINB>
__global__ void __launch_bounds__(kDiagThreads, MINB) diag_inverse_kernel(cx<T>* __restrict__ Q, cx<T>* __restrict__ Dinv,
                                                                          int np, size_t mat_elems, size_t dinv_pitch, int kb) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<T>* S = reinterpret_cast<cx<T>*>(smem_raw);
  cx<T>* scratch = S + 4096;
  cx<T>* q = Q + (size_t)blockIdx.x * mat_elems + (size_t)kb * 64 * np + (size_t)kb * 64;
  const int tid = threadIdx.x;
  if (tid < kDiagWork) {
#pragma unroll
    for (int u = 0; u < 4096 / kDiagWork; ++u) { const int e = tid + u * kDiagWork, r = e >> 6, c = e & 63; S[e] = q[(size_t)r * np + c]; }
  }
  __syncthreads();
  gj_inverse_smem<T, 64, 4, kDiagThreads>(S, scratch, [](int r, int c) { return r * 64 + c; }, [] { __syncthreads(); });
  __syncthreads();
  cx<T>* d = Dinv + (size_t)blockIdx.x * dinv_pitch;
  if (tid < kDiagWork) {
#pragma unroll
    for (int u = 0; u < 4096 / kDiagWork; ++u) {
      const int e = tid + u * kDiagWork, r = e >> 6, c = e & 63;
      const cx<T> v = S[e];
      d[e] = v;
      q[(size_t)r * np + c] = v;
    }
  }
}

template <typename T>
static gemm::Task<T> square_task(const Layout<T>& L, const cx<T>* A, const cx<T>* B, cx<T>* C) {
  gemm::Task<T> t = gemm::make_task<T>();
  t.A0 = A; t.B0 = B; t.C1 = C;
  t.sA0 = t.sB0 = t.sC1 = (long long)L.mat_elems;
  t.lda0 = t.ldb0 = t.ldc1 = L.np;
  t.M = t.N = t.K = L.np;
  return t;
}
template <typename T>
static void add_term(const Layout<T>& L, gemm::Task<T>& t, const cx<T>* A, const cx<T>* B) {
  t.A1 = A; t.B1 = B; t.sA1 = t.sB1 = (long long)L.mat_elems; t.lda1 = t.ldb1 = L.np;
}
template <typename T>
static void set_x(const Layout<T>& L, gemm::Task<T>& t, int i, const cx<T>* X, double coef) {
  if (i == 0) { t.X0 = X; t.x0 = (T)coef; t.sX0 = (long long)L.mat_elems; t.ldx0 = L.np; }
  if (i == 1) { t.X1 = X; t.x1 = (T)coef; t.sX1 = (long long)L.mat_elems; t.ldx1 = L.np; }
  if (i == 2) { t.X2 = X; t.x2 = (T)coef; t.sX2 = (long long)L.mat_elems; t.ldx2 = L.np; }
}
template <typename T>
static void set_c2(const Layout<T>& L, gemm::Task<T>& t, cx<T>* C2) { t.C2 = C2; t.sC2 = (long long)L.mat_elems; t.ldc2 = L.np; }

template <typename T>
static int run1(const Layout<T>& L, const gemm::Task<T>& t, cudaStream_t st) {
  gemm::Launch<T> G; G.t[0] = t; G.t[1] = t; G.ntasks = 1; G.batch = (int)L.batch; G.s = L.s; G.sym = L.sym;
  return gemm::launch(G, st);
}
template <typename T>
static int run2(const Layout<T>& L, const gemm::Task<T>& a, const gemm::Task<T>& b, cudaStream_t st) {
  gemm::Launch<T> G; G.t[0] = a; G.t[1] = b; G.ntasks = 2; G.batch = (int)L.batch; G.s = L.s; G.sym = L.sym;
  return gemm::launch(G, st);
}

// In-place blocked Gauss-Jordan inverse of Q (np x np per matrix).  Three launches per 64-wide step:
//   1. the diagonal block's inverse (also stored back into Q_kk);
//   2. both panels in one launch: row panel Q[k,j] <- Dinv Q[k,j] (j != k); column panel
//      Q[i,k] <- -Q[i,k] Dinv (i != k), whose second output saves the OLD column into colsave;
//   3. the rank-64 trailing update Q[i,j] -= oldcol[i] . newrow[j]  (i, j != k).
template <typename T>
static int block_inverse(const Layout<T>& L, cx<T>* Q, cudaStream_t st) {
  const int nb = L.np / 64;
  const size_t dpitch = align_up((size_t)scratch_bk<T>() * scratch_bk<T>() * sizeof(cx<T>), 256) / sizeof(cx<T>);
  const size_t cpitch = align_up((size_t)L.np * scratch_bk<T>() * sizeof(cx<T>), 256) / sizeof(cx<T>);
  const size_t smem = (4096 + 4 * 64 + 4) * sizeof(cx<T>);
  static std::atomic<unsigned long long> attr_done{0};
  if (first_use_on_device(attr_done)) {
    QG_RETURN_IF_CUDA(cudaFuncSetAttribute(diag_inverse_kernel<T, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QG_RETURN_IF_CUDA(cudaFuncSetAttribute(diag_inverse_kernel<T, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  for (int k = 0; k < nb; ++k) {
    if (L.batch > 148) diag_inverse_kernel<T, 2><<<(unsigned)L.batch, kDiagThreads, smem, st>>>(Q, L.dinv, L.np, L.mat_elems, dpitch, k);
    else diag_inverse_kernel<T, 1><<<(unsigned)L.batch, kDiagThreads, smem, st>>>(Q, L.dinv, L.np, L.mat_elems, dpitch, k);
    QG_CHECK_LAUNCH();
    if (nb == 1) break;
    gemm::Task<T> rp = gemm::make_task<T>();
    rp.A0 = L.dinv; rp.sA0 = (long long)dpitch; rp.lda0 = 64;
    rp.B0 = Q + (size_t)k * 64 * L.np; rp.sB0 = (long long)L.mat_elems; rp.ldb0 = L.np;
    rp.C1 = Q + (size_t)k * 64 * L.np; rp.sC1 = (long long)L.mat_elems; rp.ldc1 = L.np;
    rp.M = 64; rp.N = L.np; rp.K = 64; rp.skip_tn = k;
    gemm::Task<T> cp = gemm::make_task<T>();
    cp.A0 = Q + (size_t)k * 64; cp.sA0 = (long long)L.mat_elems; cp.lda0 = L.np;
    cp.B0 = L.dinv; cp.sB0 = (long long)dpitch; cp.ldb0 = 64;
    cp.C1 = Q + (size_t)k * 64; cp.sC1 = (long long)L.mat_elems; cp.ldc1 = L.np;
    cp.al1 = T(-1); cp.sg1 = T(0);
    cp.C2 = L.colsave; cp.sC2 = (long long)cpitch; cp.ldc2 = 64; cp.al2 = T(0); cp.sg2 = T(1);
    cp.X0 = Q + (size_t)k * 64; cp.sX0 = (long long)L.mat_elems; cp.ldx0 = L.np; cp.x0 = T(1);
    cp.M = L.np; cp.N = 64; cp.K = 64; cp.skip_tm = k;
    int rc = run2(L, rp, cp, st); if (rc) return rc;
    gemm::Task<T> tr = gemm::make_task<T>();
    tr.A0 = L.colsave; tr.sA0 = (long long)cpitch; tr.lda0 = 64;
    tr.B0 = Q + (size_t)k * 64 * L.np; tr.sB0 = (long long)L.mat_elems; tr.ldb0 = L.np;
    tr.C1 = Q; tr.sC1 = (long long)L.mat_elems; tr.ldc1 = L.np;
    tr.X0 = Q; tr.sX0 = (long long)L.mat_elems; tr.ldx0 = L.np; tr.x0 = T(1);
    tr.al1 = T(-1); tr.sg1 = T(1);
    tr.M = L.np; tr.N = L.np; tr.K = 64; tr.skip_tm = k; tr.skip_tn = k;
    rc = run1(L, tr, st); if (rc) return rc;
  }
  return QG_OK;
}

// ---- complex64: the same sweep in 128-wide blocks with every product on the tensor-core kernel (gemm_tc.cu) ----------------
// With the chain's products at 230+ TFLOP/s the 64-wide sweep on the FMA kernel (rank-64 updates, ~40 TFLOP/s) had become a
// quarter of a complex64 expm.  Per 128-wide step: the diagonal block's inverse in one CTA (register-blocked sweep of
// common.cuh, 8 x 8 per thread), the row panel, a copy of the OLD pivot column, the rank-128 trailing update, and last the
// column panel from the saved column (with two column tiles per row block an in-place column panel would race).
constexpr int kDiag128Work = 256, kDiag128Threads = kDiag128Work + 32;
__global__ void __launch_bounds__(kDiag128Threads, 1) diag_inverse128_kernel(cx<float>* __restrict__ Q, cx<float>* __restrict__ Dinv, int np,
                                                                             size_t mat_elems, size_t dinv_pitch, int kb) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<float>* S = reinterpret_cast<cx<float>*>(smem_raw);
  cx<float>* scratch = S + 128 * 128;
  cx<float>* q = Q + (size_t)blockIdx.x * mat_elems + (size_t)kb * 128 * np + (size_t)kb * 128;
  const int tid = threadIdx.x;
  for (int e = tid; e < 128 * 128; e += kDiag128Threads) { const int r = e >> 7, c = e & 127; S[e] = q[(size_t)r * np + c]; }
  __syncthreads();
  gj_inverse_smem<float, 128, 8, kDiag128Threads>(S, scratch, [](int r, int c) { return r * 128 + c; }, [] { __syncthreads(); });
  __syncthreads();
  cx<float>* d = Dinv + (size_t)blockIdx.x * dinv_pitch;
  for (int e = tid; e < 128 * 128; e += kDiag128Threads) {
    const int r = e >> 7, c = e & 127;
    const cx<float> v = S[e];
    d[e] = v;
    q[(size_t)r * np + c] = v;
  }
}
// colsave[b][i][0 .. 127] = Q[b][i][128 kb ...]   grid (np / 8, batch), block (128, 8)... one float4 = two elements per thread
__global__ void save_column128_kernel(const cx<float>* __restrict__ Q, cx<float>* __restrict__ colsave, int np, size_t mat_elems,
                                      size_t cpitch, int kb) {
  const int b = blockIdx.y;
  const int r = blockIdx.x * 4 + (threadIdx.x >> 6), c2 = (threadIdx.x & 63) * 2;
  const float4 v = *reinterpret_cast<const float4*>(Q + (size_t)b * mat_elems + (size_t)r * np + (size_t)kb * 128 + c2);
  *reinterpret_cast<float4*>(colsave + (size_t)b * cpitch + (size_t)r * 128 + c2) = v;
}
static int run_tc(const Layout<float>& L, const gemm::Task<float>& t, cudaStream_t st) {
  gemm::Launch<float> G; G.t[0] = t; G.t[1] = t; G.ntasks = 1; G.batch = (int)L.batch; G.s = L.s; G.sym = nullptr;
  const int rc = gemm::tc::launch(G, st);
  return rc == gemm::tc::kNotTaken ? QG_ERR_ARG : rc;
}
static bool block_inverse_tc_ok(const Layout<float>& L) { return gemm::tc::usable() && L.np % 128 == 0 && L.batch <= 65535; }
static int block_inverse_tc(const Layout<float>& L, cx<float>* Q, cudaStream_t st) {
  using T = float;
  const int np = L.np, nb = np / 128;
  const size_t dpitch = align_up((size_t)128 * 128 * sizeof(cx<T>), 256) / sizeof(cx<T>);
  const size_t cpitch = align_up((size_t)np * 128 * sizeof(cx<T>), 256) / sizeof(cx<T>);
  const size_t smem = (size_t)(128 * 128 + 4 * 128 + 4) * sizeof(cx<T>);
  static std::atomic<unsigned long long> attr_done{0};
  if (first_use_on_device(attr_done))
    QG_RETURN_IF_CUDA(cudaFuncSetAttribute(diag_inverse128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int k = 0; k < nb; ++k) {
    diag_inverse128_kernel<<<(unsigned)L.batch, kDiag128Threads, smem, st>>>(Q, L.dinv, np, L.mat_elems, dpitch, k);
    QG_CHECK_LAUNCH();
    if (nb == 1) break;
    int rc;
    {
      // row panel Q[k, j] <- Dinv Q[k, j], j != k (in place: a CTA reads exactly the columns it writes)
      gemm::Task<T> t = gemm::make_task<T>();
      t.A0 = L.dinv; t.sA0 = (long long)dpitch; t.lda0 = 128;
      t.B0 = Q + (size_t)k * 128 * np; t.sB0 = (long long)L.mat_elems; t.ldb0 = np;
      t.C1 = Q + (size_t)k * 128 * np; t.sC1 = (long long)L.mat_elems; t.ldc1 = np;
      t.M = 128; t.N = np; t.K = 128; t.skip_c0 = k * 128;
      rc = run_tc(L, t, st); if (rc) return rc;
    }
    save_column128_kernel<<<dim3(np / 4, (unsigned)L.batch), 256, 0, st>>>(Q, L.colsave, np, L.mat_elems, cpitch, k);
    QG_CHECK_LAUNCH();
    {
      // trailing update Q[i, j] -= oldcol[i] . newrow[j], i, j != k
      gemm::Task<T> t = gemm::make_task<T>();
      t.A0 = L.colsave; t.sA0 = (long long)cpitch; t.lda0 = 128;
      t.B0 = Q + (size_t)k * 128 * np; t.sB0 = (long long)L.mat_elems; t.ldb0 = np;
      t.C1 = Q; t.sC1 = (long long)L.mat_elems; t.ldc1 = np;
      t.X0 = Q; t.sX0 = (long long)L.mat_elems; t.ldx0 = np; t.x0 = T(1);
      t.al1 = T(-1); t.sg1 = T(1);
      t.M = np; t.N = np; t.K = 128; t.skip_r0 = k * 128; t.skip_c0 = k * 128;
      rc = run_tc(L, t, st); if (rc) return rc;
    }
    {
      // column panel Q[i, k] <- -oldcol[i] Dinv, i != k
      gemm::Task<T> t = gemm::make_task<T>();
      t.A0 = L.colsave; t.sA0 = (long long)cpitch; t.lda0 = 128;
      t.B0 = L.dinv; t.sB0 = (long long)dpitch; t.ldb0 = 128;
      t.C1 = Q + (size_t)k * 128; t.sC1 = (long long)L.mat_elems; t.ldc1 = np;
      t.al1 = T(-1); t.sg1 = T(0);
      t.M = np; t.N = 128; t.K = 128; t.skip_r0 = k * 128;
      rc = run_tc(L, t, st); if (rc) return rc;
    }
  }
  return QG_OK;
}
template <typename T> static int inverse_sweep(const Layout<T>& L, cx<T>* Q, cudaStream_t st) {
  if constexpr (std::is_same<T, float>::value) {
    if (block_inverse_tc_ok(L)) return block_inverse_tc(L, Q, st);
  }
  return block_inverse(L, Q, st);
}

// ---- Q^-1 for one or two matrices (the 8- and 4-rank shards of the learning step), complex128, order 7.
// The blocked Gauss-Jordan sweep is a chain of np/64 dependent steps of three launches each; with a single
// matrix every one of them is latency bound (1.54 ms at n = 1024, a fifth of the rank's step).  Here instead
//     X0 = sum_{k<=7} d_k As^k   (Taylor series of 1/q_7, from the powers already formed: one product)
//     Qi = X0 (2 I - Q X0)       (one Newton-Schulz step: two products)
// q_7 has no zero below |z| = 4.6 and every scaled matrix has ||As|| <= 0.783 (1-norm, or 2-norm for the normal
// ones), where the series tail sum_{k>=8} |d_k| z^k is 6.2e-8: ||I - Q X0|| <= 1e-7 and one step squares it.
// Three n^3 products at tensor-pipe speed: 0.8 ms.  Larger batches keep the sweep (its updates fill the GPU there).
template <typename T>
__global__ void odd_series_kernel(Layout<T> L, cx<T>* __restrict__ Od, T cI, T c2, T c4, T c6) {
  const int b = blockIdx.y;
  const cx<T>* A2 = L.fwd(B_A2) + (size_t)b * L.mat_elems;
  const cx<T>* A4 = L.fwd(B_A4) + (size_t)b * L.mat_elems;
  const cx<T>* A6 = L.fwd(B_A6) + (size_t)b * L.mat_elems;
  cx<T>* o = Od + (size_t)b * L.mat_elems;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < L.mat_elems; e += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / L.np), c = (int)(e % L.np);
    const cx<T> a = A2[e], bq = A4[e], cq = A6[e];
    o[e] = cx<T>(fma(c6, cq.re, fma(c4, bq.re, fma(c2, a.re, r == c ? cI : T(0)))), fma(c6, cq.im, fma(c4, bq.im, c2 * a.im)));
  }
}
template <typename T>
static int series_inverse(const Layout<T>& L, cudaStream_t st, bool keep) {
  // d_k: Taylor coefficients of 1 / q_7(z), q_7(z) = sum_k (b_k / b_0) (-z)^k
  const double d[8] = {1.0, 0.5, 7.0 / 52.0, 1.0 / 39.0, 343.0 / 89232.0, 107.0 / 223080.0, 5383.0 / 104401440.0,
                       893.0 / 182702520.0};
  cx<T>* As = L.fwd(B_AS); cx<T>* A2 = L.fwd(B_A2); cx<T>* A4 = L.fwd(B_A4); cx<T>* A6 = L.fwd(B_A6); cx<T>* Q = L.fwd(B_Q);
  cx<T>* X0 = L.R(1); cx<T>* Y = L.R(0);                              // both free until R0 = Qi.P
  const size_t gx_ = (L.mat_elems + 255) / 256; const int gx = (int)(gx_ < 1024 ? gx_ : 1024);
  odd_series_kernel<T><<<dim3(gx, (unsigned)L.batch), 256, 0, st>>>(L, Y, (T)d[1], (T)d[3], (T)d[5], (T)d[7]);
  QG_CHECK_LAUNCH();
  int rc;
  {
    // X0 = As.Od + Ev.  (Skew-)Hermitian A: Ev is Hermitian and As.Od skew-Hermitian / Hermitian, exactly the U-GEMM's
    // pattern -- with a second output Ev - As.Od only the upper-triangle tiles are computed and the mirror of X0 is the
    // conjugate of the second output.  That output is scratch: the first dual buffer, idle during the forward pass of a
    // value + pull-back call (a value-only call has no spare buffer and computes the full product).
    gemm::Task<T> t = square_task(L, As, Y, X0);
    set_x(L, t, 0, A2, d[2]); set_x(L, t, 1, A4, d[4]); set_x(L, t, 2, A6, d[6]); t.xI = (T)d[0];
    if (keep) { t.herm = 2; set_c2(L, t, L.dual(D_E)); t.al2 = T(-1); t.sg2 = T(1); }
    rc = run1(L, t, st); if (rc) return rc;
  }
  {
    gemm::Task<T> t = square_task(L, Q, X0, Y);                        // Y = 2 I - Q.X0
    t.al1 = T(-1); t.xI = T(2);
    rc = run1(L, t, st); if (rc) return rc;
  }
  return run1(L, square_task(L, X0, Y, Q), st);                        // Qi = X0.Y (over Q)
}

template <typename T>
static int forward(const void* Ag, void* Yg, int64_t batch, int n, void* ws, size_t ws_bytes, int max_sq, bool keep,
                   cudaStream_t st) {
  Layout<T> L;
  const int mode = keep ? QG_EXPM_FWD_VJP : QG_EXPM_FWD;
  if (!carve(L, ws, ws_bytes, n, batch, mode, max_sq)) return QG_ERR_WORKSPACE;
  if (batch > 16384) return QG_ERR_ARG;
  const int np = L.np;
  const int m = PadePlan<T>::kTop;
  const cx<T>* A = (const cx<T>*)Ag;
  const size_t cpitch = align_up((size_t)np * sizeof(T), 256);
  QG_RETURN_IF_CUDA(cudaMemsetAsync(L.colsum, 0, (size_t)batch * cpitch, st));
  QG_RETURN_IF_CUDA(cudaMemsetAsync(L.asym, 0, (size_t)batch * 3 * sizeof(T), st));
  {
    dim3 g((n + 63) / 64, (n + 63) / 64, (unsigned)batch);
    colsum_kernel<T><<<g, 256, 0, st>>>(A, L.colsum, L.asym, n, n, (size_t)n * n, np);
    QG_CHECK_LAUNCH();
    plan_kernel<T><<<(unsigned)batch, 256, 0, st>>>(L, keep ? 1 : 0);
    QG_CHECK_LAUNCH();
    const size_t gx_ = (L.mat_elems + 255) / 256; const int gx = (int)(gx_ < 1024 ? gx_ : 1024);
    scale_pad_kernel<T><<<dim3(gx, (unsigned)batch), 256, 0, st>>>(A, L.fwd(B_AS), L.s, n, np);
    QG_CHECK_LAUNCH();
  }
  cx<T>* As = L.fwd(B_AS); cx<T>* A2 = L.fwd(B_A2); cx<T>* A4 = L.fwd(B_A4); cx<T>* A6 = L.fwd(B_A6);
  cx<T>* W = L.fwd(B_W); cx<T>* Q = L.fwd(B_Q); cx<T>* P = L.fwd(B_P);
  int rc;
  // (skew-)Hermitian A: A2, A4, A6 (and W, V) are Hermitian -- only the upper-triangle tiles are computed
  { gemm::Task<T> t = square_task(L, As, As, A2); t.herm = 1; rc = run1(L, t, st); if (rc) return rc; }
  { gemm::Task<T> t = square_task(L, A2, A2, A4); t.herm = 1; rc = run1(L, t, st); if (rc) return rc; }
  if (m == 7) { gemm::Task<T> t = square_task(L, A4, A2, A6); t.herm = 1; rc = run1(L, t, st); if (rc) return rc; }
  {
    // s of the normal matrices from the norm of the top even power, then the late rescale and W
    QG_RETURN_IF_CUDA(cudaMemsetAsync(L.colsum, 0, (size_t)batch * cpitch, st));
    dim3 g(np / 64, np / 64, (unsigned)batch);
    colsum_kernel<T><<<g, 256, 0, st>>>(m == 7 ? A6 : A4, L.colsum, nullptr, np, np, L.mat_elems, np);
    QG_CHECK_LAUNCH();
    power_gate_kernel<T><<<(unsigned)batch, 256, 0, st>>>(L, keep ? 1 : 0, m == 7 ? 6 : 4);
    QG_CHECK_LAUNCH();
    const size_t xs_bytes = (size_t)4 * np * sizeof(cx<T>);
    const int iters = plan_rigorous() ? 0 : kPowerIters;               // rigorous mode: s from d_k alone (an upper bound on rho)
    if (iters == 0) {
      // nothing to run: plan_power_kernel below takes d_k
    } else if (batch >= 128 && np <= 128) {
      power_iter_fused_kernel<T><<<(unsigned)batch, 256, xs_bytes, st>>>(L, m == 7 ? A6 : A4, kPowerIters, keep ? 1 : 0, m == 7 ? 6 : 4);
      QG_CHECK_LAUNCH();
    } else {
      const bool few = (long long)batch * (np / 8) < 2LL * gemm::device_sm_count();   // a warp per row would leave SMs idle
      for (int it = 0; it < kPowerIters; ++it) {
        if (few) power_iter_kernel<T, 4><<<dim3(np / 2, (unsigned)batch), 256, 0, st>>>(L, m == 7 ? A6 : A4, it, keep ? 1 : 0, m == 7 ? 6 : 4);
        else power_iter_kernel<T, 1><<<dim3(np / 8, (unsigned)batch), 256, 0, st>>>(L, m == 7 ? A6 : A4, it, keep ? 1 : 0, m == 7 ? 6 : 4);
        QG_CHECK_LAUNCH();
      }
    }
    plan_power_kernel<T><<<(unsigned)batch, 256, 0, st>>>(L, keep ? 1 : 0, m == 7 ? 6 : 4, iters);
    QG_CHECK_LAUNCH();
    const size_t gx_ = (L.mat_elems + 255) / 256; const int gx = (int)(gx_ < 1024 ? gx_ : 1024);
    rescale_w_kernel<T><<<dim3(gx, (unsigned)batch), 256, 0, st>>>(L, m, (T)pade_coef(m, 1), (T)pade_coef(m, 3),
                                                                 (T)pade_coef(m, 5), (T)pade_coef(m, 7));
    QG_CHECK_LAUNCH();
  }
  {
    // U = As.W;  Q = V - U,  P = V + U,  V = c6 A6 + c4 A4 + c2 A2 + c0 I
    // ((skew-)Hermitian A: U = A W is skew-Hermitian / Hermitian, V Hermitian -- Q and P mirror each other / themselves)
    gemm::Task<T> t = square_task(L, As, W, Q);
    t.herm = 2;
    set_c2(L, t, P); t.al1 = T(-1); t.sg1 = T(1); t.al2 = T(1); t.sg2 = T(1);
    set_x(L, t, 0, A2, pade_coef(m, 2)); set_x(L, t, 1, A4, pade_coef(m, 4));
    if (m == 7) set_x(L, t, 2, A6, pade_coef(m, 6));
    t.xI = (T)pade_coef(m, 0);
    rc = run1(L, t, st); if (rc) return rc;
  }
  rc = (sizeof(T) == 8 && m == 7 && batch <= 2) ? series_inverse(L, st, keep) : inverse_sweep(L, Q, st);
  if (rc) return rc;
  rc = run1(L, square_task(L, Q, P, L.R(0)), st); if (rc) return rc;
  for (int it = 0; it < max_sq; ++it) {
    gemm::Task<T> t = square_task(L, L.R(it), L.R(it), L.R(it + 1));
    t.pred_it = it;
    rc = run1(L, t, st); if (rc) return rc;
  }
  if (Yg) {
    const size_t gx_ = ((size_t)n * n + 255) / 256; const int gx = (int)(gx_ < 1024 ? gx_ : 1024);
    finalize_kernel<T><<<dim3(gx, (unsigned)batch), 256, 0, st>>>(L, (cx<T>*)Yg);
    QG_CHECK_LAUNCH();
  }
  return QG_OK;
}

template <typename T>
static int backward(const void* Ybar, void* Abar, int64_t batch, int n, int adjoint, void* ws, size_t ws_bytes,
                    int max_sq, cudaStream_t st) {
  Layout<T> L;
  if (!carve(L, ws, ws_bytes, n, batch, QG_EXPM_FWD_VJP, max_sq)) return QG_ERR_WORKSPACE;
  if (batch > 16384) return QG_ERR_ARG;
  const int np = L.np;
  const int m = PadePlan<T>::kTop;
  const int conj_adj = adjoint == QG_ADJ_CONJ ? 1 : 0;
  cx<T>* As = L.fwd(B_AS); cx<T>* A2 = L.fwd(B_A2); cx<T>* A4 = L.fwd(B_A4);
  cx<T>* W = L.fwd(B_W); cx<T>* Qi = L.fwd(B_Q);
  cx<T>* E = L.dual(D_E); cx<T>* M2 = L.dual(D_M2); cx<T>* M4 = L.dual(D_M4); cx<T>* M6 = L.dual(D_M6);
  cx<T>* Lw = L.dual(D_LW);
  {
    dim3 g(np / 32, np / 32, (unsigned)batch);
    scale_pad_adjoint_kernel<T><<<g, dim3(32, 8), 0, st>>>((const cx<T>*)Ybar, E, L.s, n, np, conj_adj, 1);
    QG_CHECK_LAUNCH();
  }
  int rc;
  { gemm::Task<T> t = square_task(L, As, E, M2); add_term(L, t, E, As); rc = run1(L, t, st); if (rc) return rc; }
  if (m == 7) {
    { gemm::Task<T> t = square_task(L, A2, M2, M4); add_term(L, t, M2, A2); rc = run1(L, t, st); if (rc) return rc; }
    // M6 = A4.M2 + M4.A2, second output Lw = c7 M6 + c5 M4 + c3 M2
    gemm::Task<T> t = square_task(L, A4, M2, M6); add_term(L, t, M4, A2);
    t.sg1 = T(0); set_c2(L, t, Lw); t.al2 = (T)pade_coef(7, 7); t.sg2 = T(1);
    set_x(L, t, 0, M4, pade_coef(7, 5)); set_x(L, t, 1, M2, pade_coef(7, 3));
    rc = run1(L, t, st); if (rc) return rc;
  } else {
    gemm::Task<T> t = square_task(L, A2, M2, M4); add_term(L, t, M2, A2);
    t.sg1 = T(0); set_c2(L, t, Lw); t.al2 = (T)pade_coef(5, 5); t.sg2 = T(1);
    set_x(L, t, 0, M2, pade_coef(5, 3));
    rc = run1(L, t, st); if (rc) return rc;
  }
  {
    // Lu = As.Lw + E.W;  D1 = Lu + Lv -> M2 (in place),  D2 = Lu - Lv -> M4 (in place)
    gemm::Task<T> t = square_task(L, As, Lw, M2); add_term(L, t, E, W);
    set_c2(L, t, M4); t.al1 = T(1); t.sg1 = T(1); t.al2 = T(1); t.sg2 = T(-1);
    set_x(L, t, 0, M2, pade_coef(m, 2)); set_x(L, t, 1, M4, pade_coef(m, 4));
    if (m == 7) set_x(L, t, 2, M6, pade_coef(m, 6));
    rc = run1(L, t, st); if (rc) return rc;
  }
  {
    // G = D2.R0 + D1 -> M6
    gemm::Task<T> t = square_task(L, M4, L.R(0), M6);
    set_x(L, t, 0, M2, 1.0);
    rc = run1(L, t, st); if (rc) return rc;
  }
  // L0 = Qi.G -> Lw ; L ping-pongs between Lw (even) and E (odd)
  rc = run1(L, square_task(L, Qi, M6, Lw), st); if (rc) return rc;
  for (int it = 0; it < max_sq; ++it) {
    cx<T>* Lc = (it & 1) ? E : Lw;
    cx<T>* Ln = (it & 1) ? Lw : E;
    gemm::Task<T> t = square_task(L, L.R(it), Lc, Ln); add_term(L, t, Lc, L.R(it));
    t.pred_it = it;
    rc = run1(L, t, st); if (rc) return rc;
  }
  {
    dim3 g((n + 31) / 32, (n + 31) / 32, (unsigned)batch);
    finalize_adjoint_kernel<T><<<g, dim3(32, 8), 0, st>>>(L, (cx<T>*)Abar, conj_adj);
    QG_CHECK_LAUNCH();
  }
  return QG_OK;
}

// Pull-back for skew-Hermitian A (U = exp(A) unitary) when only the SKEW-HERMITIAN part of Abar is wanted -- which
// is all a real parameter of a Hermitian generator can see: <Abar, dA/dtheta> with dA/dtheta skew-Hermitian.
// Input Z = Ybar U^H (the cotangent in the body frame; the learning step forms it as G Phi^H without ever building
// Ybar).  With Zs = (Z - Z^H)/2:
//     skew(Abar) = S(1),   S(h) = int_0^h e^{-xA} Zs e^{xA} dx,   S(2h) = S(h) + e^{-hA} S(h) e^{hA}
//     S(2^-s) = R0^H L(As, 2^-s Zs) = Qi^H (D1 + D2 R0)           (R0^H Qi = Qi^H: Q P^-1 = R0^-1, P = Q^H)
// Every direction in the Frechet recurrences is now skew-Hermitian next to a skew-Hermitian As, so M2, M4, M6, Lw, Lv
// are Hermitian, Lu is skew-Hermitian (D2' := Lv - Lu = D1^H) and every S is skew-Hermitian: those products compute
// the upper-triangle tiles only and mirror them.  GEMM-equivalents: 4 x 2t + 1 + t + s (1 + t), t = (T+1)/2T
// (T = np/64; 0.53 at n = 1024), against 10 + 2 s for the general pull-back: 12.4 vs 18 at s = 4.
// Matrices that are not skew-Hermitian (sym != 1) come back NaN.
template <typename T>
static int backward_skew(const void* Zg, void* Abar, int64_t batch, int n, void* ws, size_t ws_bytes, int max_sq,
                         cudaStream_t st) {
  Layout<T> L;
  if (!carve(L, ws, ws_bytes, n, batch, QG_EXPM_FWD_VJP, max_sq)) return QG_ERR_WORKSPACE;
  if (batch > 16384) return QG_ERR_ARG;
  const int np = L.np;
  const int m = PadePlan<T>::kTop;
  cx<T>* As = L.fwd(B_AS); cx<T>* A2 = L.fwd(B_A2); cx<T>* A4 = L.fwd(B_A4);
  cx<T>* W = L.fwd(B_W); cx<T>* Qi = L.fwd(B_Q);
  cx<T>* E = L.dual(D_E); cx<T>* M2 = L.dual(D_M2); cx<T>* M4 = L.dual(D_M4); cx<T>* M6 = L.dual(D_M6);
  cx<T>* Lw = L.dual(D_LW);
  const dim3 gt(np / 32, np / 32, (unsigned)batch), bt(32, 8);
  skew_pad_kernel<T><<<gt, bt, 0, st>>>((const cx<T>*)Zg, E, L.s, n, np);
  QG_CHECK_LAUNCH();
  int rc;
  { gemm::Task<T> t = square_task(L, As, E, M2); add_term(L, t, E, As); t.herm = 1; rc = run1(L, t, st); if (rc) return rc; }
  if (m == 7) {
    { gemm::Task<T> t = square_task(L, A2, M2, M4); add_term(L, t, M2, A2); t.herm = 1; rc = run1(L, t, st); if (rc) return rc; }
    gemm::Task<T> t = square_task(L, A4, M2, M6); add_term(L, t, M4, A2); t.herm = 1;
    t.sg1 = T(0); set_c2(L, t, Lw); t.al2 = (T)pade_coef(7, 7); t.sg2 = T(1);
    set_x(L, t, 0, M4, pade_coef(7, 5)); set_x(L, t, 1, M2, pade_coef(7, 3));
    rc = run1(L, t, st); if (rc) return rc;
  } else {
    gemm::Task<T> t = square_task(L, A2, M2, M4); add_term(L, t, M2, A2); t.herm = 1;
    t.sg1 = T(0); set_c2(L, t, Lw); t.al2 = (T)pade_coef(5, 5); t.sg2 = T(1);
    set_x(L, t, 0, M2, pade_coef(5, 3));
    rc = run1(L, t, st); if (rc) return rc;
  }
  {
    // Lu = As.Lw + E.W (skew-Hermitian), Lv Hermitian:  D1 = Lu + Lv -> M2,  D2' = Lv - Lu = D1^H -> M4 (pair mirror)
    gemm::Task<T> t = square_task(L, As, Lw, M2); add_term(L, t, E, W); t.herm = 2;
    set_c2(L, t, M4); t.al1 = T(1); t.sg1 = T(1); t.al2 = T(-1); t.sg2 = T(1);
    set_x(L, t, 0, M2, pade_coef(m, 2)); set_x(L, t, 1, M4, pade_coef(m, 4));
    if (m == 7) set_x(L, t, 2, M6, pade_coef(m, 6));
    rc = run1(L, t, st); if (rc) return rc;
  }
  {
    // G = D1 - D2'.R0 -> M6
    gemm::Task<T> t = square_task(L, M4, L.R(0), M6);
    t.al1 = T(-1); set_x(L, t, 0, M2, 1.0);
    rc = run1(L, t, st); if (rc) return rc;
  }
  // S0 = Qi^H.G -> E  (Qi^H materialised in Lw)
  scale_pad_adjoint_kernel<T><<<gt, bt, 0, st>>>(Qi, Lw, L.s, np, np, 1, 0);
  QG_CHECK_LAUNCH();
  { gemm::Task<T> t = square_task(L, Lw, M6, E); t.herm = 3; rc = run1(L, t, st); if (rc) return rc; }
  // S_{k+1} = S_k + R_k^H (S_k R_k):  S ping-pongs between E (even) and M2 (odd), T = S_k R_k in M6, R_k^H in Lw
  for (int it = 0; it < max_sq; ++it) {
    cx<T>* Sc = (it & 1) ? M2 : E;
    cx<T>* Sn = (it & 1) ? E : M2;
    scale_pad_adjoint_kernel<T><<<gt, bt, 0, st>>>(L.R(it), Lw, L.s, np, np, 1, 0);
    QG_CHECK_LAUNCH();
    { gemm::Task<T> t = square_task(L, Sc, L.R(it), M6); t.pred_it = it; rc = run1(L, t, st); if (rc) return rc; }
    gemm::Task<T> t = square_task(L, Lw, M6, Sn); set_x(L, t, 0, Sc, 1.0); t.herm = 3; t.pred_it = it;
    rc = run1(L, t, st); if (rc) return rc;
  }
  {
    const size_t gx_ = ((size_t)n * n + 255) / 256; const int gx = (int)(gx_ < 1024 ? gx_ : 1024);
    finalize_skew_kernel<T><<<dim3(gx, (unsigned)batch), 256, 0, st>>>(L, (cx<T>*)Abar);
    QG_CHECK_LAUNCH();
  }
  return QG_OK;
}

}  // namespace large

int expm_set_plan_rigorous(int on) {
  const bool prev = large::plan_rigorous();
  large::g_plan_rigorous.store(on ? 1 : 0, std::memory_order_relaxed);
  return prev ? 1 : 0;
}

size_t expm_large_workspace_bytes(int dtype, int n, int64_t batch, int mode, int max_sq) {
  return dtype == QG_C128 ? large::workspace_bytes<double>(n, batch, mode, max_sq)
                          : large::workspace_bytes<float>(n, batch, mode, max_sq);
}
int expm_large_fwd(int dtype, const void* A, void* Y, int64_t batch, int n, void* ws, size_t ws_bytes, int max_sq,
                   bool keep, cudaStream_t st) {
  return dtype == QG_C128 ? large::forward<double>(A, Y, batch, n, ws, ws_bytes, max_sq, keep, st)
                          : large::forward<float>(A, Y, batch, n, ws, ws_bytes, max_sq, keep, st);
}
int expm_large_bwd(int dtype, const void* Ybar, void* Abar, int64_t batch, int n, int adjoint, void* ws,
                   size_t ws_bytes, int max_sq, cudaStream_t st) {
  if (adjoint == QG_ADJ_SKEW_BODY)
    return dtype == QG_C128 ? large::backward_skew<double>(Ybar, Abar, batch, n, ws, ws_bytes, max_sq, st)
                            : large::backward_skew<float>(Ybar, Abar, batch, n, ws, ws_bytes, max_sq, st);
  return dtype == QG_C128 ? large::backward<double>(Ybar, Abar, batch, n, adjoint, ws, ws_bytes, max_sq, st)
                          : large::backward<float>(Ybar, Abar, batch, n, adjoint, ws, ws_bytes, max_sq, st);
}

// Generic strided-batched C = A.B used by qg_gemm (ket products of the learning step).
int gemm_plain(int dtype, const void* A, const void* B, void* C, int64_t batch, int m, int n, int k, int64_t lda,
               int64_t ldb, int64_t ldc, int64_t sA, int64_t sB, int64_t sC, cudaStream_t st) {
  if (k % 16 || batch > 65535) return QG_ERR_DIM;
  if (dtype == QG_C64 && !(lda % 2 || ldb % 2)) {
    // the tensor-core kernel takes ragged shapes (TMA zero-fills the overhang, the epilogue is predicated)
    gemm::Task<float> t = gemm::make_task<float>();
    t.A0 = (const cx<float>*)A; t.B0 = (const cx<float>*)B; t.C1 = (cx<float>*)C;
    t.sA0 = sA; t.sB0 = sB; t.sC1 = sC; t.lda0 = (int)lda; t.ldb0 = (int)ldb; t.ldc1 = (int)ldc; t.M = m; t.N = n; t.K = k;
    gemm::Launch<float> G; G.t[0] = t; G.t[1] = t; G.ntasks = 1; G.batch = (int)batch; G.s = nullptr;
    const int rc = gemm::tc::launch(G, st);
    if (rc != gemm::tc::kNotTaken) return rc;
  }
  if (m % 64 || n % 64) return QG_ERR_DIM;
  if (dtype == QG_C128) {
    gemm::Task<double> t = gemm::make_task<double>();
    t.A0 = (const cx<double>*)A; t.B0 = (const cx<double>*)B; t.C1 = (cx<double>*)C;
    t.sA0 = sA; t.sB0 = sB; t.sC1 = sC; t.lda0 = (int)lda; t.ldb0 = (int)ldb; t.ldc1 = (int)ldc; t.M = m; t.N = n; t.K = k;
    gemm::Launch<double> G; G.t[0] = t; G.t[1] = t; G.ntasks = 1; G.batch = (int)batch; G.s = nullptr;
    return gemm::launch(G, st);
  }
  if (lda % 2 || ldb % 2) return QG_ERR_DIM;
  gemm::Task<float> t = gemm::make_task<float>();
  t.A0 = (const cx<float>*)A; t.B0 = (const cx<float>*)B; t.C1 = (cx<float>*)C;
  t.sA0 = sA; t.sB0 = sB; t.sC1 = sC; t.lda0 = (int)lda; t.ldb0 = (int)ldb; t.ldc1 = (int)ldc; t.M = m; t.N = n; t.K = k;
  gemm::Launch<float> G; G.t[0] = t; G.t[1] = t; G.ntasks = 1; G.batch = (int)batch; G.s = nullptr;
  return gemm::launch(G, st);
}

}  // namespace qg

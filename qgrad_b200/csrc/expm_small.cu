// K1 / K3 -- small-dimension (5 <= n <= 32; complex64: 9 <= n <= 32) batched complex matrix exponential and
// its reverse-mode pull-back, one fused kernel, every intermediate resident in shared memory.  (Smaller n
// go to the row-per-thread kernel of expm_tiny.cu; the packed 8 x 8 slots below remain for completeness.)
//
// Replaces scipy.linalg.expm as called by squeeze (qgrad/qgrad_qutip.py:266-280) and make_unitary
// (examples/Unitary-Learning-no-qgrad.py:112-141), and the finite-difference gradient
// (Unitary-Learning-no-qgrad.py:213-241) by the Frechet derivative.
//
// One CTA owns one NP x NP slot, NP in {8, 16, 32} (n is zero-padded up to NP; for n <= 4 and
// n <= 2 two / four matrices are packed block-diagonally into one 8 x 8 slot).  complex128
// products run on the FP64 tensor pipe (mma.sync m8n8k4 = SASS DMMA.8x8x4, four real MMAs per
// complex 8x8x4 step); complex64 uses FP32 FMAs.  Shared-memory matrices are stored with an XOR
// swizzle on the column index so that both the A-fragment (2 rows x 64 B per 8 lanes) and the
// B-fragment (4 rows x 32 B per 8 lanes) 16-byte loads are bank-conflict free.
#include <stdlib.h>
#include "common.cuh"
#ifndef QG_NW16F
#define QG_NW16F 2
#endif
#ifndef QG_NW32F
#define QG_NW32F 8
#endif
#ifndef QG_NW16D
#define QG_NW16D 2
#endif

namespace qg {
namespace small {

template <int NP> __device__ __forceinline__ int sidx(int r, int c) {
  return r * NP + (c ^ (((r & 1) << 2) | (r & 2)));
}

template <int NW> __device__ __forceinline__ void cta_sync() {
  if (NW == 1) __syncwarp(); else __syncthreads();
}

template <typename T>
struct MMArgs {
  const cx<T>* A0 = nullptr; const cx<T>* B0 = nullptr;   // acc = A0*B0 (+ A1*B1)
  const cx<T>* A1 = nullptr; const cx<T>* B1 = nullptr;
  cx<T>* C1 = nullptr; cx<T>* C2 = nullptr;               // C1 = al1*acc + sg1*lin, C2 = al2*acc + sg2*lin
  T al1 = 1, sg1 = 1, al2 = 1, sg2 = 1;
  const cx<T>* X0 = nullptr; const cx<T>* X1 = nullptr; const cx<T>* X2 = nullptr;
  T x0 = 0, x1 = 0, x2 = 0, xI = 0;                       // lin = x0*X0 + x1*X1 + x2*X2 + xI*I
  bool sync_store = false;                                // CTA barrier between the products and the stores: lets C1 / C2
                                                          // overwrite an OPERAND of this very product (every warp of the CTA
                                                          // must then call mm() with the same flag)
};

template <typename T, int NP>
__device__ __forceinline__ void mm_store(const MMArgs<T>& g, int r, int c, cx<T> acc) {
  const int i = sidx<NP>(r, c);
  cx<T> lin(r == c ? g.xI : T(0), T(0));
  if (g.X0) { cx<T> x = g.X0[i]; lin.re = fma(g.x0, x.re, lin.re); lin.im = fma(g.x0, x.im, lin.im); }
  if (g.X1) { cx<T> x = g.X1[i]; lin.re = fma(g.x1, x.re, lin.re); lin.im = fma(g.x1, x.im, lin.im); }
  if (g.X2) { cx<T> x = g.X2[i]; lin.re = fma(g.x2, x.re, lin.re); lin.im = fma(g.x2, x.im, lin.im); }
  cx<T> o1(fma(g.al1, acc.re, g.sg1 * lin.re), fma(g.al1, acc.im, g.sg1 * lin.im));
  cx<T> o2(fma(g.al2, acc.re, g.sg2 * lin.re), fma(g.al2, acc.im, g.sg2 * lin.im));
  g.C1[i] = o1;
  if (g.C2) g.C2[i] = o2;
}

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// complex128: each warp owns TPW adjacent 8x8 output tiles of one tile row.
// (warps W0 .. W0 + NW - 1 take part; the default is the whole CTA)
template <int NP, int NW, int W0 = 0>
__device__ __forceinline__ void mm(const MMArgs<double>& g) {
  constexpr int TR = NP / 8;
  constexpr int TPW = (TR * TR) / NW;
  static_assert(TPW >= 1 && TPW * NW == TR * TR, "tile split");
  const int warp = (threadIdx.x >> 5) - W0, lane = threadIdx.x & 31;
  const int gid = lane >> 2, tig = lane & 3;
  const int ti = (warp * TPW) / TR, tj0 = (warp * TPW) % TR;
  double cr[TPW][2], ci[TPW][2];
#pragma unroll
  for (int t = 0; t < TPW; ++t) { cr[t][0] = cr[t][1] = ci[t][0] = ci[t][1] = 0.0; }
#pragma unroll
  for (int term = 0; term < 2; ++term) {
    const cx<double>* A = term ? g.A1 : g.A0;
    const cx<double>* B = term ? g.B1 : g.B0;
    if (A != nullptr) {
#pragma unroll
      for (int kk = 0; kk < NP / 4; ++kk) {
        const cx<double> a = A[sidx<NP>(ti * 8 + gid, kk * 4 + tig)];
        const double nai = -a.im;
#pragma unroll
        for (int t = 0; t < TPW; ++t) {
          const cx<double> b = B[sidx<NP>(kk * 4 + tig, (tj0 + t) * 8 + gid)];
          dmma(cr[t], a.re, b.re);
          dmma(ci[t], a.re, b.im);
          dmma(cr[t], nai, b.im);
          dmma(ci[t], a.im, b.re);
        }
      }
    }
  }
  if (g.sync_store) __syncthreads();
#pragma unroll
  for (int t = 0; t < TPW; ++t)
#pragma unroll
    for (int e = 0; e < 2; ++e)
      mm_store<double, NP>(g, ti * 8 + gid, (tj0 + t) * 8 + tig * 2 + e, cx<double>(cr[t][e], ci[t][e]));
}

// (overlap helper: only the complex128 kernel ever takes this path)
template <int NP, int NWS, int W0> __device__ __forceinline__ void mm_lu(const MMArgs<double>& g) { mm<NP, NWS, W0>(g); }
template <int NP, int NWS, int W0> __device__ __forceinline__ void mm_lu(const MMArgs<float>&) {}

// complex64: FP32 FMAs, each thread a TM x 2 strip.  (An earlier version of this product, with single-element
// loads and a partially unrolled k loop, only gave stable results as a __noinline__ function; this one is inlined
// and covered by tools/stress_c64_small.py: 20000-matrix batches, bit-identical across runs, every matrix within
// tolerance of the complex128 kernel.)
template <int NP, int NW>
__device__ __forceinline__ void mm(const MMArgs<float>& g) {
  constexpr int NT = NW * 32;
  constexpr int TM = (NP * NP) / (2 * NT);
  static_assert(TM >= 1, "strip");
  const int rp = threadIdx.x / (NP / 2), c = 2 * (threadIdx.x % (NP / 2));
  cx<float> acc[TM][2];
#pragma unroll
  for (int i = 0; i < TM; ++i) { acc[i][0] = cx<float>(0.f, 0.f); acc[i][1] = cx<float>(0.f, 0.f); }
  // the swizzle XORs even values into the column index, so the element pairs (k, k+1) of a row of A and
  // (c, c+1) of a row of B are adjacent: one 16-byte load each, 2 + TM loads per 8 TM complex MACs
#pragma unroll
  for (int term = 0; term < 2; ++term) {
    const cx<float>* A = term ? g.A1 : g.A0;
    const cx<float>* B = term ? g.B1 : g.B0;
    if (A != nullptr) {
#pragma unroll 4
      for (int k = 0; k < NP; k += 2) {
        const float4 b0 = *reinterpret_cast<const float4*>(&B[sidx<NP>(k, c)]);
        const float4 b1 = *reinterpret_cast<const float4*>(&B[sidx<NP>(k + 1, c)]);
        const cx<float> b00(b0.x, b0.y), b01(b0.z, b0.w), b10(b1.x, b1.y), b11(b1.z, b1.w);
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          const float4 a = *reinterpret_cast<const float4*>(&A[sidx<NP>(rp * TM + i, k)]);
          const cx<float> a0(a.x, a.y), a1(a.z, a.w);
          cfma(acc[i][0], a0, b00); cfma(acc[i][1], a0, b01);
          cfma(acc[i][0], a1, b10); cfma(acc[i][1], a1, b11);
        }
      }
    }
  }
  if (g.sync_store) cta_sync<NW>();
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    mm_store<float, NP>(g, rp * TM + i, c, acc[i][0]);
    mm_store<float, NP>(g, rp * TM + i, c + 1, acc[i][1]);
  }
}

// One-warp slots (NP = 8): the plain in-place sweep, two elements per lane; the register-blocked
// gj_inverse_smem of common.cuh (used for NP >= 16) needs more than one warp to pay off.
template <typename T, int NP>
__device__ __forceinline__ void gj_inverse_warp(cx<T>* Q) {
  constexpr int EPT = (NP * NP) / 32;
  for (int k = 0; k < NP; ++k) {
    const cx<T> pinv = crecip(Q[sidx<NP>(k, k)]);
    cx<T> nv[EPT];
#pragma unroll
    for (int q = 0; q < EPT; ++q) {
      const int e = threadIdx.x + q * 32, r = e / NP, c = e % NP;
      const cx<T> v = Q[sidx<NP>(r, c)];
      if (r == k) {
        nv[q] = (c == k) ? pinv : v * pinv;
      } else {
        const cx<T> f = Q[sidx<NP>(r, k)] * pinv;
        if (c == k) nv[q] = -f;
        else { const cx<T> gkc = Q[sidx<NP>(k, c)]; nv[q] = v - f * gkc; }
      }
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < EPT; ++q) {
      const int e = threadIdx.x + q * 32;
      Q[sidx<NP>(e / NP, e % NP)] = nv[q];
    }
    __syncwarp();
  }
}

constexpr int kNormalGain = 6;     // most squarings a normal matrix can save: ||A||_1 <= sqrt(n) ||A||_2

// max column abs-sum of an NP x NP shared-memory matrix, NaN-propagating; every thread gets it.
// red: NT + 1 scratch values.  Ends with a CTA barrier.
template <typename T, int NP, int NW>
__device__ __forceinline__ T colmax(const cx<T>* M, T* red) {
  constexpr int NT = NW * 32;
  const int tid = threadIdx.x;
  const int c = tid % NP, rg = tid / NP;
  T part = T(0);
  for (int r = rg; r < NP; r += NT / NP) part += cabs_norm(M[sidx<NP>(r, c)]);
  red[tid] = part;
  cta_sync<NW>();
  if (tid < 32) {
    T best = T(0);
    for (int cc = tid; cc < NP; cc += 32) {
      T s = T(0);
      for (int q = 0; q < NT / NP; ++q) s += red[q * NP + cc];
      best = (s > best || s != s) ? s : best;               // NaN propagates
    }
    for (int o = 16; o > 0; o >>= 1) {
      T other = __shfl_xor_sync(0xffffffffu, best, o);
      best = (other > best || other != other) ? other : best;
    }
    if (tid == 0) red[NT] = best;
  }
  cta_sync<NW>();
  const T v = red[NT];
  cta_sync<NW>();
  return v;
}

// warps per CTA (= per matrix slot).  complex64 at NP = 16 runs ONE warp per matrix: the FP32 products are
// short, so warp-level barriers (no CTA barrier anywhere) and 8 outputs per lane beat four warps that meet
// at ~25 CTA barriers per matrix.
template <typename T, int NP> struct Cfg { static constexpr int NW = NP == 8 ? 1 : (NP == 16 ? (sizeof(T) == 4 ? QG_NW16F : QG_NW16D) : (sizeof(T) == 4 ? QG_NW32F : 8)); };

#ifdef QG_DEBUG
__device__ int d_dbg_stage = -1;
#define QG_DUMP(stage, bufptr)                                                                   \
  if (d_dbg_stage == (stage)) {                                                                  \
    cta_sync<NW>();                                                                              \
    for (int e = tid; e < NN; e += NT) {                                                         \
      const int r = e / NP, c = e % NP;                                                          \
      if (r < n && c < n && mat0 < batch) Yg[(size_t)mat0 * nn + (size_t)r * n + c] = (bufptr)[sidx<NP>(r, c)]; \
    }                                                                                            \
    return;                                                                                      \
  }
#else
#define QG_DUMP(stage, bufptr)
#endif

template <typename T, int NP, bool VJP>
__global__ void __launch_bounds__(Cfg<T, NP>::NW * 32)
expm_small_kernel(const cx<T>* __restrict__ Ag, const cx<T>* __restrict__ Ybar,
                  cx<T>* __restrict__ Yg, cx<T>* __restrict__ Abar,
                  long long batch, int n, int sub, int conj_adjoint) {
  constexpr int NW = Cfg<T, NP>::NW;
  constexpr int NT = NW * 32;
  constexpr int NN = NP * NP;
  constexpr int NBUF = VJP ? 10 : 5;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<T>* buf = reinterpret_cast<cx<T>*>(smem_raw);
  T* red = reinterpret_cast<T*>(buf + NBUF * NN);          // scratch: NT + 4 reals (norms) / 4 NP + 4 complex (inverse)
  cx<T>* bA = buf;
  cx<T>* bA2 = buf + 1 * NN; cx<T>* bA4 = buf + 2 * NN; cx<T>* bA6 = buf + 3 * NN; cx<T>* bW = buf + 4 * NN;
  cx<T>* bE = buf + 5 * NN; cx<T>* bM2 = buf + 6 * NN; cx<T>* bM4 = buf + 7 * NN; cx<T>* bM6 = buf + 8 * NN;
  cx<T>* bLw = buf + 9 * NN;   // (the last five exist only when VJP)
  const int tid = threadIdx.x;
  const int G = NP / sub;                                   // matrices packed block-diagonally
  const long long mat0 = (long long)blockIdx.x * G;
  const size_t nn = (size_t)n * n;

  // ---- load A (zero padded / block diagonal) and its 1-norm; the cotangent is fetched in the same
  //      (coalesced) order right away, into registers, so its latency hides under the norm and the plan
  constexpr int EPT = NN / NT;
  cx<T> vy[VJP ? EPT : 1];
#pragma unroll
  for (int q = 0; q < EPT; ++q) {
    const int e = tid + q * NT;
    const int r = e / NP, c = e % NP, g = r / sub, rl = r % sub, cl = c - g * sub;
    cx<T> v(T(0), T(0)), y(T(0), T(0));
    if (cl >= 0 && cl < n && rl < n && mat0 + g < batch) {
      const size_t o = (size_t)(mat0 + g) * nn + (size_t)rl * n + cl;
      v = Ag[o];
      if (VJP) y = Ybar[o];
    }
    bA[sidx<NP>(r, c)] = v;
    if (VJP) vy[q] = y;
  }
  cta_sync<NW>();
  const T n1 = colmax<T, NP, NW>(bA, red);
  int m, s;
  pade_plan<T>((double)n1, VJP, &m, &s);
  const bool bad = s >= 64;                                 // inf / NaN norm
  if (bad) s = 0;
  // A (skew-)Hermitian slot is normal: rho(A) = ||A||_2 <= ||A^k||_1^(1/k), so the number of squarings
  // can come from the norm of the top even power (computed below anyway) instead of ||A||_1; such a
  // slot is pre-scaled kNormalGain levels less and settles its s after the powers (expm_large.cu has
  // the long version of this comment).
  bool normal = false;
  // (8 x 8 slots skip the test: ||A||_1 <= sqrt(8) ||A||_2 leaves at most one level to gain, less than the test costs)
  if (NP >= 16 && s > 0) {
    T d_skew = T(0), d_herm = T(0), tot = T(0);
    for (int e = tid; e < NN; e += NT) {
      const int r = e / NP, c = e % NP;
      const cx<T> v = bA[sidx<NP>(r, c)], w = bA[sidx<NP>(c, r)];
      d_skew += cabs2_(cx<T>(v.re + w.re, v.im - w.im));      // squared distances: no square roots needed for the test
      d_herm += cabs2_(cx<T>(v.re - w.re, v.im + w.im));
      tot += cabs2_(v);
    }
    d_skew = warp_sum(d_skew); d_herm = warp_sum(d_herm); tot = warp_sum(tot);
    if ((tid & 31) == 0) { red[(tid >> 5) * 3] = d_skew; red[(tid >> 5) * 3 + 1] = d_herm; red[(tid >> 5) * 3 + 2] = tot; }
    cta_sync<NW>();
    d_skew = d_herm = tot = T(0);
    for (int w = 0; w < NW; ++w) { d_skew += red[w * 3]; d_herm += red[w * 3 + 1]; tot += red[w * 3 + 2]; }
    cta_sync<NW>();
    const T tol2 = sizeof(T) == 8 ? T(8.3e-25) : T(1.5e-11);  // (2^-40)^2, (2^-18)^2: Frobenius distance relative to ||A||_F
    normal = fmin(d_skew, d_herm) <= tol2 * tot;
  }
  if (normal) s = s > kNormalGain ? s - kNormalGain : 0;
  const T scale = (T)ldexp(1.0, -s);

  // ---- scale A; E = Ybar^H (or Ybar^T) scaled: each in-block element goes to its transposed place
#pragma unroll
  for (int q = 0; q < EPT; ++q) {
    const int e = tid + q * NT;
    const int r = e / NP, c = e % NP, i = sidx<NP>(r, c);
    bA[i] = scale * bA[i];
    if (VJP) {
      const int g = r / sub, rl = r % sub, cl = c - g * sub;
      cx<T> v = vy[q];
      if (conj_adjoint) v.im = -v.im;
      const bool inblock = cl >= 0 && cl < sub;
      bE[inblock ? sidx<NP>(g * sub + cl, g * sub + rl) : i] = scale * v;
    }
  }
  cta_sync<NW>();

  QG_DUMP(0, bA)
  const T c0 = (T)pade_coef(m, 0), c1 = (T)pade_coef(m, 1), c2 = (T)pade_coef(m, 2), c3 = (T)pade_coef(m, 3),
          c4 = (T)pade_coef(m, 4), c5 = (T)pade_coef(m, 5), c6 = (T)pade_coef(m, 6), c7 = (T)pade_coef(m, 7);
  const bool has4 = m >= 5, has6 = m >= 7;

  // ---- even powers (and their duals)
  { MMArgs<T> g; g.A0 = bA; g.B0 = bA; g.C1 = bA2; mm<NP, NW>(g); }
  if (VJP) { MMArgs<T> g; g.A0 = bA; g.B0 = bE; g.A1 = bE; g.B1 = bA; g.C1 = bM2; mm<NP, NW>(g); }
  cta_sync<NW>();
  QG_DUMP(1, bA2)
  if (has4) {
    { MMArgs<T> g; g.A0 = bA2; g.B0 = bA2; g.C1 = bA4; mm<NP, NW>(g); }
    if (VJP) { MMArgs<T> g; g.A0 = bA2; g.B0 = bM2; g.A1 = bM2; g.B1 = bA2; g.C1 = bM4; mm<NP, NW>(g); }
    cta_sync<NW>();
  }
  if (has6) {
    { MMArgs<T> g; g.A0 = bA4; g.B0 = bA2; g.C1 = bA6; mm<NP, NW>(g); }
    if (VJP) { MMArgs<T> g; g.A0 = bA4; g.B0 = bM2; g.A1 = bM4; g.B1 = bA2; g.C1 = bM6; mm<NP, NW>(g); }
    cta_sync<NW>();
  }
  QG_DUMP(2, bA4)
  // ---- normal slots: e extra levels from d_k = ||A^k||_1^(1/k) (k = 6, or 4 at order 5), applied as
  //      A *= 2^-e, A2 *= 2^-2e, A4 *= 2^-4e, A6 *= 2^-6e (same factors for E, M2, M4, M6) in the loop below
  int ext = 0;
  if (normal) {
    const T pk = colmax<T, NP, NW>(has6 ? bA6 : bA4, red);
    const double dk = pow((double)pk, has6 ? 1.0 / 6.0 : 0.25);
    const double n1s = ldexp((double)n1, -s);
    double x = (dk < n1s) ? dk : n1s;
    if (pk != pk) x = n1s;
    const double bound = PadePlan<T>::bound(PadePlan<T>::kTop, VJP);
    while (!(x <= bound) && ext < 64) { x *= 0.5; ++ext; }
    s += ext;
  }
  const T f1 = (T)ldexp(1.0, -ext), f2 = (T)ldexp(1.0, -2 * ext), f4 = (T)ldexp(1.0, -4 * ext), f6 = (T)ldexp(1.0, -6 * ext);
  // ---- W = c7 A6 + c5 A4 + c3 A2 + c1 I ,  Lw = c7 M6 + c5 M4 + c3 M2   (physical-index loop)
  for (int e = tid; e < NN; e += NT) {
    const int r = e / NP, c = e % NP, i = sidx<NP>(r, c);
    cx<T> w(r == c ? c1 : T(0), T(0));
    if (ext) { bA[i] = f1 * bA[i]; bA2[i] = f2 * bA2[i]; if (has4) bA4[i] = f4 * bA4[i]; if (has6) bA6[i] = f6 * bA6[i]; }
    { cx<T> x = bA2[i]; w.re = fma(c3, x.re, w.re); w.im = fma(c3, x.im, w.im); }
    if (has4) { cx<T> x = bA4[i]; w.re = fma(c5, x.re, w.re); w.im = fma(c5, x.im, w.im); }
    if (has6) { cx<T> x = bA6[i]; w.re = fma(c7, x.re, w.re); w.im = fma(c7, x.im, w.im); }
    bW[i] = w;
    if (VJP) {
      if (ext) { bE[i] = f1 * bE[i]; bM2[i] = f2 * bM2[i]; if (has4) bM4[i] = f4 * bM4[i]; if (has6) bM6[i] = f6 * bM6[i]; }
      cx<T> x = bM2[i];
      cx<T> lw(c3 * x.re, c3 * x.im);
      if (has4) { x = bM4[i]; lw.re = fma(c5, x.re, lw.re); lw.im = fma(c5, x.im, lw.im); }
      if (has6) { x = bM6[i]; lw.re = fma(c7, x.re, lw.re); lw.im = fma(c7, x.im, lw.im); }
      bLw[i] = lw;
    }
  }
  cta_sync<NW>();
  QG_DUMP(3, bW)
  // ---- U = A W;  Q = V - U -> bA2,  P = V + U -> bA4,  V = c6 A6 + c4 A4 + c2 A2 + c0 I
  {
    MMArgs<T> g; g.A0 = bA; g.B0 = bW; g.C1 = bA2; g.C2 = bA4;
    g.al1 = T(-1); g.sg1 = T(1); g.al2 = T(1); g.sg2 = T(1);
    g.X0 = bA2; g.x0 = c2; if (has4) { g.X1 = bA4; g.x1 = c4; } if (has6) { g.X2 = bA6; g.x2 = c6; }
    g.xI = c0;
    mm<NP, NW>(g);
  }
  // ---- Lu = A Lw + E W;  D1 = Lu + Lv -> bM2,  D2 = Lu - Lv -> bM4,  Lv = c6 M6 + c4 M4 + c2 M2
  //      complex128 with >= 4 warps: the Gauss-Jordan sweep below keeps only (NP/B)^2 / 32 + 1 warps busy and
  //      does not touch anything Lu reads or writes, so the remaining warps compute Lu UNDER the sweep.
  constexpr int GJ_B = NP >= 32 ? 4 : 2;
  constexpr int GJ_WARPS = ((NP / GJ_B) * (NP / GJ_B) + 31) / 32 + 1;
  constexpr int LU_WARPS = NW - GJ_WARPS >= 4 ? 4 : 1;
  constexpr bool OVERLAP = VJP && sizeof(T) == 8 && LU_WARPS == 4 && (NP / 8) * (NP / 8) / LU_WARPS <= NP / 8;   // NP = 32
  MMArgs<T> glu;
  if (VJP) {
    glu.A0 = bA; glu.B0 = bLw; glu.A1 = bE; glu.B1 = bW; glu.C1 = bM2; glu.C2 = bM4;
    glu.al1 = T(1); glu.sg1 = T(1); glu.al2 = T(1); glu.sg2 = T(-1);
    glu.X0 = bM2; glu.x0 = c2; if (has4) { glu.X1 = bM4; glu.x1 = c4; } if (has6) { glu.X2 = bM6; glu.x2 = c6; }
    if (!OVERLAP) mm<NP, NW>(glu);
  }
  cta_sync<NW>();
  // ---- Qi = Q^-1 (in place),  R = Qi P,  G = D1 + D2 R,  L = Qi G
  QG_DUMP(4, bA2)
  QG_DUMP(5, bA4)
  if constexpr (NW == 1) {
    gj_inverse_warp<T, NP>(bA2);
  } else if constexpr (OVERLAP) {
    const int warp_id = tid >> 5;
    if (warp_id < GJ_WARPS) {
      gj_inverse_smem<T, NP, GJ_B, GJ_WARPS * 32>(bA2, reinterpret_cast<cx<T>*>(red), [](int r, int c) { return sidx<NP>(r, c); },
                                                   [] { asm volatile("bar.sync 1, %0;" ::"n"(GJ_WARPS * 32) : "memory"); });
    } else if (warp_id >= NW - LU_WARPS) {
      mm_lu<NP, LU_WARPS, NW - LU_WARPS>(glu);
    }
    cta_sync<NW>();
  } else {
    gj_inverse_smem<T, NP, (NP >= 32 ? 4 : 2), NT>(bA2, reinterpret_cast<cx<T>*>(red), [](int r, int c) { return sidx<NP>(r, c); },
                                                    [] { cta_sync<NW>(); });
    cta_sync<NW>();
  }
  QG_DUMP(6, bA2)
  { MMArgs<T> g; g.A0 = bA2; g.B0 = bA4; g.C1 = bA6; mm<NP, NW>(g); }
  cta_sync<NW>();
  QG_DUMP(7, bA6)
  cx<T>* R = bA6; cx<T>* Rn = bW;
  cx<T>* L = nullptr; cx<T>* Ln = nullptr;
  if (VJP) {
    { MMArgs<T> g; g.A0 = bM4; g.B0 = bA6; g.C1 = bM6; g.X0 = bM2; g.x0 = T(1); mm<NP, NW>(g); }
    cta_sync<NW>();
    { MMArgs<T> g; g.A0 = bA2; g.B0 = bM6; g.C1 = bLw; mm<NP, NW>(g); }
    cta_sync<NW>();
    L = bLw; Ln = bE; Rn = bA;
  }
  // ---- squarings: L <- R L + L R,  R <- R R
  for (int it = 0; it < s; ++it) {
    { MMArgs<T> g; g.A0 = R; g.B0 = R; g.C1 = Rn; mm<NP, NW>(g); }
    if (VJP) { MMArgs<T> g; g.A0 = R; g.B0 = L; g.A1 = L; g.B1 = R; g.C1 = Ln; mm<NP, NW>(g); }
    cta_sync<NW>();
    cx<T>* t = R; R = Rn; Rn = t;
    if (VJP) { t = L; L = Ln; Ln = t; }
  }
  // ---- store Y = R and Abar = L^H (or L^T)
  const T nanv = bad ? (T)nan("") : T(0);
  for (int e = tid; e < NN; e += NT) {
    const int r = e / NP, c = e % NP, g = r / sub, rl = r % sub, cl = c - g * sub;
    if (cl >= 0 && cl < n && rl < n && mat0 + g < batch) {
      const size_t o = (size_t)(mat0 + g) * nn + (size_t)rl * n + cl;
      if (Yg != nullptr) { cx<T> v = R[sidx<NP>(r, c)]; v.re += nanv; v.im += nanv; Yg[o] = v; }
      if (VJP) {
        cx<T> v = L[sidx<NP>(g * sub + cl, g * sub + rl)];    // transposed read
        if (conj_adjoint) v.im = -v.im;
        v.re += nanv; v.im += nanv;
        Abar[o] = v;
      }
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Fused value + pull-back in SIX shared-memory buffers (NP = 16, 32) instead of the ten of the kernel above.
// At NP = 32 in complex128 ten buffers are 160 KB: one CTA per SM, and ncu shows the FP64 tensor pipe idle 55 % of the
// time -- every barrier, the norm passes, the load / store phases and the (FMA-pipe, latency-bound) Gauss-Jordan sweep
// leave it with nothing to do.  Six buffers are 96 KB: TWO CTAs per SM, whose product phases fill each other's gaps.
// Slot schedule (S0 .. S5; "->" = written into; every product's inputs are dead or rewritten only after a barrier):
//   load            A -> S0, E = Ybar^H -> S1 (scaled)
//   1               A2 = A A -> S2                    M2 = A E + E A -> S3
//   2  (m >= 5)     A4 = A2 A2 -> S4                  M4 = A2 M2 + M2 A2 -> S5
//   3  (m == 7)     A6 = A4 A2 -> S0 (A is dead until U)   M6 = A4 M2 + M4 A2 -> S1 (E likewise)
//   4  elementwise  m == 7: W -> S0 (over A6), V -> S4 (over A4), Lw -> S1 (over M6), Lv -> S5 (over M4);
//                           then A and E are RE-READ from global memory (they are in L2: +50 % input traffic on a
//                           kernel that moves 2 % of the HBM roofline) -> S2, S3
//                   m == 5: W -> S2 (over A2), V -> S4, Lw -> S3 (over M2), Lv -> S5; A, E still in S0, S1
//                   m == 3: W -> S4, V -> S2 (in place), Lw -> S5, Lv -> S3 (in place); A, E still in S0, S1
//   5               Lu = A Lw + E W:  D1 = Lu + Lv -> slot of Lw (barrier before the stores),  D2 = Lu - Lv -> slot of Lv
//   6               U = A W:          Q = V - U -> slot of W (barrier before the stores),       P = V + U -> slot of V
//   7               Qi = Q^-1 in place
//   8 .. 10         R = Qi P -> slot of A;   G = D1 + D2 R -> slot of E;   L = Qi G -> slot of P
//   squarings       R, L ping-pong with the slots of D1, D2
template <typename T, int NP>
__global__ void __launch_bounds__(Cfg<T, NP>::NW * 32, (NP == 32 && sizeof(T) == 8) ? 2 : 1)
expm_small_vjp6_kernel(const cx<T>* __restrict__ Ag, const cx<T>* __restrict__ Ybar, cx<T>* __restrict__ Yg,
                       cx<T>* __restrict__ Abar, long long batch, int n, int conj_adjoint) {
  constexpr int NW = Cfg<T, NP>::NW;
  constexpr int NT = NW * 32;
  constexpr int NN = NP * NP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<T>* buf = reinterpret_cast<cx<T>*>(smem_raw);
  T* red = reinterpret_cast<T*>(buf + 6 * NN);
  cx<T>* S0 = buf; cx<T>* S1 = buf + NN; cx<T>* S2 = buf + 2 * NN; cx<T>* S3 = buf + 3 * NN; cx<T>* S4 = buf + 4 * NN; cx<T>* S5 = buf + 5 * NN;
  const int tid = threadIdx.x;
  const long long mat = blockIdx.x;
  const size_t nn = (size_t)n * n;
  const cx<T>* Am = Ag + (size_t)mat * nn;
  const cx<T>* Ym = Ybar + (size_t)mat * nn;

  // A (zero padded) -> S0, scaled once s is known; E = Ybar^H (or ^T) -> S1
  auto load_inputs = [&](cx<T>* dA, cx<T>* dE, T scale) {
    for (int e = tid; e < NN; e += NT) {
      const int r = e / NP, c = e % NP;
      cx<T> v(T(0), T(0)), y(T(0), T(0));
      if (r < n && c < n) { v = Am[(size_t)r * n + c]; y = Ym[(size_t)r * n + c]; }
      if (conj_adjoint) y.im = -y.im;
      dA[sidx<NP>(r, c)] = scale * v;
      dE[sidx<NP>(c, r)] = scale * y;
    }
  };
  load_inputs(S0, S1, T(1));
  cta_sync<NW>();
  const T n1 = colmax<T, NP, NW>(S0, red);
  int m, s;
  pade_plan<T>((double)n1, true, &m, &s);
  const bool bad = s >= 64;                                 // inf / NaN norm
  if (bad) s = 0;
  bool normal = false;
  if (s > 0) {                                              // (skew-)Hermitian slot -> the 2-norm rule (see the kernel above)
    T d_skew = T(0), d_herm = T(0), tot = T(0);
    for (int e = tid; e < NN; e += NT) {
      const int r = e / NP, c = e % NP;
      const cx<T> v = S0[sidx<NP>(r, c)], w = S0[sidx<NP>(c, r)];
      d_skew += cabs2_(cx<T>(v.re + w.re, v.im - w.im));
      d_herm += cabs2_(cx<T>(v.re - w.re, v.im + w.im));
      tot += cabs2_(v);
    }
    d_skew = warp_sum(d_skew); d_herm = warp_sum(d_herm); tot = warp_sum(tot);
    if ((tid & 31) == 0) { red[(tid >> 5) * 3] = d_skew; red[(tid >> 5) * 3 + 1] = d_herm; red[(tid >> 5) * 3 + 2] = tot; }
    cta_sync<NW>();
    d_skew = d_herm = tot = T(0);
    for (int w = 0; w < NW; ++w) { d_skew += red[w * 3]; d_herm += red[w * 3 + 1]; tot += red[w * 3 + 2]; }
    cta_sync<NW>();
    const T tol2 = sizeof(T) == 8 ? T(8.3e-25) : T(1.5e-11);
    normal = fmin(d_skew, d_herm) <= tol2 * tot;
  }
  if (normal) s = s > kNormalGain ? s - kNormalGain : 0;
  if (s > 0) {
    const T scale = (T)ldexp(1.0, -s);
    for (int e = tid; e < NN; e += NT) { S0[e] = scale * S0[e]; S1[e] = scale * S1[e]; }
    cta_sync<NW>();
  }
  const T c0 = (T)pade_coef(m, 0), c1 = (T)pade_coef(m, 1), c2 = (T)pade_coef(m, 2), c3 = (T)pade_coef(m, 3),
          c4 = (T)pade_coef(m, 4), c5 = (T)pade_coef(m, 5), c6 = (T)pade_coef(m, 6), c7 = (T)pade_coef(m, 7);
  const bool has4 = m >= 5, has6 = m >= 7;

  // ---- 1-3: even powers and their duals
  { MMArgs<T> g; g.A0 = S0; g.B0 = S0; g.C1 = S2; mm<NP, NW>(g); }
  { MMArgs<T> g; g.A0 = S0; g.B0 = S1; g.A1 = S1; g.B1 = S0; g.C1 = S3; mm<NP, NW>(g); }
  cta_sync<NW>();
  if (has4) {
    { MMArgs<T> g; g.A0 = S2; g.B0 = S2; g.C1 = S4; mm<NP, NW>(g); }
    { MMArgs<T> g; g.A0 = S2; g.B0 = S3; g.A1 = S3; g.B1 = S2; g.C1 = S5; mm<NP, NW>(g); }
    cta_sync<NW>();
  }
  if (has6) {
    { MMArgs<T> g; g.A0 = S4; g.B0 = S2; g.C1 = S0; mm<NP, NW>(g); }
    { MMArgs<T> g; g.A0 = S4; g.B0 = S3; g.A1 = S5; g.B1 = S2; g.C1 = S1; mm<NP, NW>(g); }
    cta_sync<NW>();
  }
  // ---- normal slots: e extra levels from the norm of the top even power
  int ext = 0;
  if (normal) {
    const T pk = colmax<T, NP, NW>(has6 ? S0 : S4, red);
    const double dk = pow((double)pk, has6 ? 1.0 / 6.0 : 0.25);
    const double n1s = ldexp((double)n1, -s);
    double x = (dk < n1s) ? dk : n1s;
    if (pk != pk) x = n1s;
    const double bound = PadePlan<T>::bound(PadePlan<T>::kTop, true);
    while (!(x <= bound) && ext < 64) { x *= 0.5; ++ext; }
    s += ext;
  }
  const T f1 = (T)ldexp(1.0, -ext), f2 = (T)ldexp(1.0, -2 * ext), f4 = (T)ldexp(1.0, -4 * ext), f6 = (T)ldexp(1.0, -6 * ext);
  // ---- 4: W, V, Lw, Lv elementwise, in place (slot table in the header comment)
  cx<T>* pA = S0; cx<T>* pE = S1; cx<T>* pW; cx<T>* pV; cx<T>* pLw; cx<T>* pLv;
  if (has6) { pW = S0; pV = S4; pLw = S1; pLv = S5; pA = S2; pE = S3; }
  else if (has4) { pW = S2; pV = S4; pLw = S3; pLv = S5; }
  else { pW = S4; pV = S2; pLw = S5; pLv = S3; }
  for (int e = tid; e < NN; e += NT) {
    const int r = e / NP, c = e % NP, i = sidx<NP>(r, c);
    const T dg = r == c ? T(1) : T(0);
    const cx<T> a2 = f2 * S2[i], m2 = f2 * S3[i];
    cx<T> a4(T(0), T(0)), m4 = a4, a6 = a4, m6 = a4;
    if (has4) { a4 = f4 * S4[i]; m4 = f4 * S5[i]; }
    if (has6) { a6 = f6 * S0[i]; m6 = f6 * S1[i]; }
    const cx<T> w(fma(c7, a6.re, fma(c5, a4.re, fma(c3, a2.re, c1 * dg))), fma(c7, a6.im, fma(c5, a4.im, c3 * a2.im)));
    const cx<T> v(fma(c6, a6.re, fma(c4, a4.re, fma(c2, a2.re, c0 * dg))), fma(c6, a6.im, fma(c4, a4.im, c2 * a2.im)));
    const cx<T> lw(fma(c7, m6.re, fma(c5, m4.re, c3 * m2.re)), fma(c7, m6.im, fma(c5, m4.im, c3 * m2.im)));
    const cx<T> lv(fma(c6, m6.re, fma(c4, m4.re, c2 * m2.re)), fma(c6, m6.im, fma(c4, m4.im, c2 * m2.im)));
    pW[i] = w; pV[i] = v; pLw[i] = lw; pLv[i] = lv;
    if (!has6 && ext) { S0[i] = f1 * S0[i]; S1[i] = f1 * S1[i]; }          // resident A, E take the late levels here
  }
  if (has6) {
    cta_sync<NW>();                                                        // S2, S3 (A2, M2) are free only now
    load_inputs(S2, S3, (T)ldexp(1.0, -s));                                // s already includes the late levels
  }
  cta_sync<NW>();
  // ---- 5: Lu = A Lw + E W;  D1 = Lu + Lv -> slot of Lw,  D2 = Lu - Lv -> slot of Lv
  {
    MMArgs<T> g; g.A0 = pA; g.B0 = pLw; g.A1 = pE; g.B1 = pW; g.C1 = pLw; g.C2 = pLv;
    g.al1 = T(1); g.sg1 = T(1); g.al2 = T(1); g.sg2 = T(-1); g.X0 = pLv; g.x0 = T(1); g.sync_store = true;
    mm<NP, NW>(g);
  }
  cta_sync<NW>();
  cx<T>* D1 = pLw; cx<T>* D2 = pLv;
  // ---- 6: U = A W;  Q = V - U -> slot of W,  P = V + U -> slot of V
  {
    MMArgs<T> g; g.A0 = pA; g.B0 = pW; g.C1 = pW; g.C2 = pV;
    g.al1 = T(-1); g.sg1 = T(1); g.al2 = T(1); g.sg2 = T(1); g.X0 = pV; g.x0 = T(1); g.sync_store = true;
    mm<NP, NW>(g);
  }
  cta_sync<NW>();
  cx<T>* Q = pW; cx<T>* P = pV;
  // ---- 7: Qi = Q^-1 in place (unpivoted: the Pade denominator is diagonally dominant)
  gj_inverse_smem<T, NP, (NP >= 32 ? 4 : 2), NT>(Q, reinterpret_cast<cx<T>*>(red), [](int r, int c) { return sidx<NP>(r, c); },
                                                  [] { cta_sync<NW>(); });
  cta_sync<NW>();
  // ---- 8-10: R = Qi P -> slot of A;  G = D1 + D2 R -> slot of E;  L = Qi G -> slot of P
  { MMArgs<T> g; g.A0 = Q; g.B0 = P; g.C1 = pA; mm<NP, NW>(g); }
  cta_sync<NW>();
  { MMArgs<T> g; g.A0 = D2; g.B0 = pA; g.C1 = pE; g.X0 = D1; g.x0 = T(1); mm<NP, NW>(g); }
  cta_sync<NW>();
  { MMArgs<T> g; g.A0 = Q; g.B0 = pE; g.C1 = P; mm<NP, NW>(g); }
  cta_sync<NW>();
  cx<T>* R = pA; cx<T>* Rn = D1; cx<T>* L = P; cx<T>* Ln = D2;
  // ---- squarings: L <- R L + L R,  R <- R R
  for (int it = 0; it < s; ++it) {
    { MMArgs<T> g; g.A0 = R; g.B0 = R; g.C1 = Rn; mm<NP, NW>(g); }
    { MMArgs<T> g; g.A0 = R; g.B0 = L; g.A1 = L; g.B1 = R; g.C1 = Ln; mm<NP, NW>(g); }
    cta_sync<NW>();
    cx<T>* t = R; R = Rn; Rn = t;
    t = L; L = Ln; Ln = t;
  }
  // ---- store Y = R and Abar = L^H (or L^T)
  const T nanv = bad ? (T)nan("") : T(0);
  for (int e = tid; e < NN; e += NT) {
    const int r = e / NP, c = e % NP;
    if (r < n && c < n) {
      const size_t o = (size_t)mat * nn + (size_t)r * n + c;
      if (Yg != nullptr) { cx<T> v = R[sidx<NP>(r, c)]; v.re += nanv; v.im += nanv; Yg[o] = v; }
      cx<T> v = L[sidx<NP>(c, r)];                                          // transposed read
      if (conj_adjoint) v.im = -v.im;
      v.re += nanv; v.im += nanv;
      Abar[o] = v;
    }
  }
}

template <typename T, int NP>
static int launch_vjp6(const void* A, const void* Ybar, void* Y, void* Abar, int64_t batch, int n, int adjoint, cudaStream_t st) {
  constexpr int NW = Cfg<T, NP>::NW;
  const size_t scratch = (size_t)(NW * 32 + 4) * sizeof(T) > (size_t)(4 * NP + 4) * sizeof(cx<T>) ? (size_t)(NW * 32 + 4) * sizeof(T)
                                                                                                   : (size_t)(4 * NP + 4) * sizeof(cx<T>);
  const size_t smem = (size_t)6 * NP * NP * sizeof(cx<T>) + scratch;
  auto kern = expm_small_vjp6_kernel<T, NP>;
  static std::atomic<unsigned long long> attr_done{0};
  if (first_use_on_device(attr_done)) {
    QG_RETURN_IF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  kern<<<(unsigned)batch, NW * 32, smem, st>>>((const cx<T>*)A, (const cx<T>*)Ybar, (cx<T>*)Y, (cx<T>*)Abar, (long long)batch, n,
                                               adjoint == QG_ADJ_CONJ ? 1 : 0);
  QG_CHECK_LAUNCH();
  return QG_OK;
}

template <typename T, int NP, bool VJP>
static int launch(const void* A, const void* Ybar, void* Y, void* Abar, int64_t batch, int n, int adjoint,
                  cudaStream_t st) {
  constexpr int NW = Cfg<T, NP>::NW;
  constexpr int NBUF = VJP ? 10 : 5;
  const size_t scratch = (size_t)(NW * 32 + 4) * sizeof(T) > (size_t)(4 * NP + 4) * sizeof(cx<T>) ? (size_t)(NW * 32 + 4) * sizeof(T)
                                                                                                   : (size_t)(4 * NP + 4) * sizeof(cx<T>);
  const size_t smem = (size_t)NBUF * NP * NP * sizeof(cx<T>) + scratch;
  const int sub = (NP == 8) ? (n <= 2 ? 2 : (n <= 4 ? 4 : 8)) : NP;
  const int G = NP / sub;
  const int64_t grid = (batch + G - 1) / G;
  auto kern = expm_small_kernel<T, NP, VJP>;
  static std::atomic<unsigned long long> attr_done{0};
  if (first_use_on_device(attr_done)) {
    QG_RETURN_IF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  // grid.x limit is 2^31-1: fine for any batch that fits in HBM
  kern<<<(unsigned)grid, NW * 32, smem, st>>>((const cx<T>*)A, (const cx<T>*)Ybar, (cx<T>*)Y, (cx<T>*)Abar,
                                              (long long)batch, n, sub, adjoint == QG_ADJ_CONJ ? 1 : 0);
  QG_CHECK_LAUNCH();
  return QG_OK;
}

template <typename T, bool VJP>
static int dispatch(const void* A, const void* Ybar, void* Y, void* Abar, int64_t batch, int n, int adjoint,
                    cudaStream_t st) {
  if (n <= 8) return launch<T, 8, VJP>(A, Ybar, Y, Abar, batch, n, adjoint, st);
  static const bool ten = [] { const char* e = getenv("QG_SMALL_VJP10"); return e && *e == '1'; }();   // A/B timing aid
  // (the six-buffer kernel at NP = 8 -- 6 KB instead of 10 KB per matrix -- is register-bound at 19 warps per SM and
  //  re-reads its inputs: measured 1.6-1.8e8 matrices/s against 2.1e8 for the ten-buffer kernel, so n <= 8 stays there)
  if (VJP && !ten) {    if (n <= 16) return launch_vjp6<T, 16>(A, Ybar, Y, Abar, batch, n, adjoint, st);
    return launch_vjp6<T, 32>(A, Ybar, Y, Abar, batch, n, adjoint, st);
  }
  if (n <= 16) return launch<T, 16, VJP>(A, Ybar, Y, Abar, batch, n, adjoint, st);
  return launch<T, 32, VJP>(A, Ybar, Y, Abar, batch, n, adjoint, st);
}

}  // namespace small

#ifdef QG_DEBUG
extern "C" int qg_debug_set_stage(int stage) {
  return (int)cudaMemcpyToSymbol(small::d_dbg_stage, &stage, sizeof(int));
}
#endif

int expm_tiny(int dtype, const void* A, const void* Ybar, void* Y, void* Abar, int64_t batch, int n, int adjoint, bool vjp,
              cudaStream_t st);   // expm_tiny.cu: n <= 4 (complex64: n <= 8), one thread per matrix row

// Entry used by api.cu: n <= 32.
int expm_small(int dtype, const void* A, const void* Ybar, void* Y, void* Abar, int64_t batch, int n,
               int adjoint, bool vjp, cudaStream_t st) {
  if (n <= 4 || (n <= 8 && dtype == QG_C64)) return expm_tiny(dtype, A, Ybar, Y, Abar, batch, n, adjoint, vjp, st);
  if (dtype == QG_C128)
    return vjp ? small::dispatch<double, true>(A, Ybar, Y, Abar, batch, n, adjoint, st)
               : small::dispatch<double, false>(A, Ybar, Y, Abar, batch, n, adjoint, st);
  return vjp ? small::dispatch<float, true>(A, Ybar, Y, Abar, batch, n, adjoint, st)
             : small::dispatch<float, false>(A, Ybar, Y, Abar, batch, n, adjoint, st);
}

}  // namespace qg

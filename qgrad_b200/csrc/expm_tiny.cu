// K1 / K3 for n <= 4 (complex64: n <= 8): batched complex matrix exponential and its reverse-mode pull-back
// with NT = 4 (or 8) THREADS PER MATRIX (thread = row), 32 / NT matrices per warp, plain FMAs.
//
// Replaces scipy.linalg.expm at the sizes the reference actually calls it with: make_unitary's 2 x 2
// blocks (examples/Unitary-Learning-no-qgrad.py:112-141) and squeeze's 4 x 4 (qgrad/qgrad_qutip.py:266-280).
// An 8 x 8 tensor-core tile would be three quarters padding at n = 4, so this kernel keeps every row in
// registers and does the 64 complex MACs of a 4 x 4 product as 16 per thread; the right-hand operand
// of a product is read from a shared-memory copy laid out [row][col][matrix] so that the eight
// matrices of a warp read one contiguous 128-byte line per (row, col) -- a single wavefront, broadcast
// within each quad -- and only matrices that ARE right-hand operands (A, E, A2, M2, W, Lw, P, R, G, L)
// are ever staged, four buffers deep.  Same algorithm as expm_small.cu (orders 3/5/7 by ||A||_1, s
// squarings, unpivoted Gauss-Jordan for the Pade denominator); orders below the top one are evaluated
// with the same instruction stream and zero coefficients, squarings beyond a matrix's own s are
// computed and discarded, so a warp never diverges.
#include "common.cuh"

namespace qg {
namespace tiny {

constexpr int kWarps = 4;                      // warps per CTA: 32 matrices
template <typename T, int NT> struct Row { cx<T> v[NT]; };

// element (k, j) of group q's matrix in buffer b (elements of cx<T>); q fastest (32 / NT matrices per warp)
template <int NT> __device__ __forceinline__ int sm_idx(int b, int k, int j, int q) { return ((b * NT + k) * NT + j) * (32 / NT) + q; }

template <typename T, int NT> __device__ __forceinline__ void stage(cx<T>* sm, int b, int row, int q, const Row<T, NT>& x) {
#pragma unroll
  for (int j = 0; j < NT; ++j) sm[sm_idx<NT>(b, row, j, q)] = x.v[j];
}
// out (+)= x . Y,  Y = buffer b of this group
template <typename T, int NT, bool ACC> __device__ __forceinline__ void mul(const cx<T>* sm, int b, int q, const Row<T, NT>& x, Row<T, NT>& out) {
  if (!ACC) {
#pragma unroll
    for (int j = 0; j < NT; ++j) out.v[j] = cx<T>(T(0), T(0));
  }
#pragma unroll
  for (int k = 0; k < NT; ++k)
#pragma unroll
    for (int j = 0; j < NT; ++j) cfma(out.v[j], x.v[k], sm[sm_idx<NT>(b, k, j, q)]);
}
template <typename T, int NT> __device__ __forceinline__ T quad_sum(T v) {
#pragma unroll
  for (int o = 1; o < NT; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <typename T> __device__ __forceinline__ cx<T> quad_bcast(cx<T> v, int src_lane) {
  return cx<T>(__shfl_sync(0xffffffffu, v.re, src_lane), __shfl_sync(0xffffffffu, v.im, src_lane));
}

template <typename T, int NT, bool VJP>
__global__ void __launch_bounds__(kWarps * 32)
expm_tiny_kernel(const cx<T>* __restrict__ Ag, const cx<T>* __restrict__ Ybar, cx<T>* __restrict__ Yg,
                 cx<T>* __restrict__ Abar, long long batch, int n, int conj_adjoint) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int MPW = 32 / NT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, q = lane / NT, row = lane % NT;
  cx<T>* sm = reinterpret_cast<cx<T>*>(smem_raw) + (size_t)warp * (VJP ? 4 : 2) * (NT * NT * MPW);
  const long long mat = ((long long)blockIdx.x * kWarps + warp) * MPW + q;
  const bool live = mat < batch;
  const size_t nn = (size_t)n * n;
  const int qbase = lane - row;

  // ---- load row `row` of A and of E = Ybar^H (or Ybar^T): zero padded to NT x NT
  Row<T, NT> a, e;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    a.v[j] = cx<T>(T(0), T(0)); e.v[j] = cx<T>(T(0), T(0));
    if (live && row < n && j < n) {
      a.v[j] = Ag[(size_t)mat * nn + (size_t)row * n + j];
      if (VJP) { cx<T> y = Ybar[(size_t)mat * nn + (size_t)j * n + row]; if (conj_adjoint) y.im = -y.im; e.v[j] = y; }
    }
  }
  // ---- ||A||_1: column sums across the quad, NaN-propagating max
  T n1 = T(0);
  bool isnan_ = false;
#pragma unroll
  for (int j = 0; j < NT; ++j) { const T cs = quad_sum<T, NT>(cabs_norm(a.v[j])); if (cs != cs) isnan_ = true; n1 = fmax(n1, cs); }
  if (isnan_) n1 = (T)INFINITY;
  int m, s;
  pade_plan<T>((double)n1, VJP, &m, &s);
  const bool bad = s >= 64;
  if (bad) s = 0;
  int smax = s;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) smax = max(smax, __shfl_xor_sync(0xffffffffu, smax, o));
  const T scale = (T)ldexp(1.0, -s);
#pragma unroll
  for (int j = 0; j < NT; ++j) { a.v[j] = scale * a.v[j]; if (VJP) e.v[j] = scale * e.v[j]; }
  const T c0 = (T)pade_coef(m, 0), c1 = (T)pade_coef(m, 1), c2 = (T)pade_coef(m, 2), c3 = (T)pade_coef(m, 3),
          c4 = (T)pade_coef(m, 4), c5 = (T)pade_coef(m, 5), c6 = (T)pade_coef(m, 6), c7 = (T)pade_coef(m, 7);

  // buffers: 0 A -> P, 1 E -> R, 2 A2 -> W -> G, 3 M2 -> Lw -> L   (forward only: 0 A -> P, 1 A2 -> W -> R)
  constexpr int BA = 0, BE = 1, BA2 = VJP ? 2 : 1, BM2 = 3;
  stage(sm, BA, row, q, a);
  if (VJP) stage(sm, BE, row, q, e);
  __syncwarp();
  Row<T, NT> a2, a4, a6, m2, m4, m6, t;
  mul<T, NT, false>(sm, BA, q, a, a2);
  if (VJP) { mul<T, NT, false>(sm, BE, q, a, m2); mul<T, NT, true>(sm, BA, q, e, m2); }
  stage(sm, BA2, row, q, a2);
  if (VJP) stage(sm, BM2, row, q, m2);
  __syncwarp();
  mul<T, NT, false>(sm, BA2, q, a2, a4);
  mul<T, NT, false>(sm, BA2, q, a4, a6);
  if (VJP) {
    mul<T, NT, false>(sm, BM2, q, a2, m4); mul<T, NT, true>(sm, BA2, q, m2, m4);
    mul<T, NT, false>(sm, BM2, q, a4, m6); mul<T, NT, true>(sm, BA2, q, m4, m6);
  }
  __syncwarp();
  // W = c7 A6 + c5 A4 + c3 A2 + c1 I, V = c6 A6 + c4 A4 + c2 A2 + c0 I (and their duals)
  Row<T, NT> w, v, lw, lv;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const T d1 = (j == row) ? c1 : T(0), d0 = (j == row) ? c0 : T(0);
    w.v[j] = cx<T>(fma(c7, a6.v[j].re, fma(c5, a4.v[j].re, fma(c3, a2.v[j].re, d1))), fma(c7, a6.v[j].im, fma(c5, a4.v[j].im, c3 * a2.v[j].im)));
    v.v[j] = cx<T>(fma(c6, a6.v[j].re, fma(c4, a4.v[j].re, fma(c2, a2.v[j].re, d0))), fma(c6, a6.v[j].im, fma(c4, a4.v[j].im, c2 * a2.v[j].im)));
    if (VJP) {
      lw.v[j] = cx<T>(fma(c7, m6.v[j].re, fma(c5, m4.v[j].re, c3 * m2.v[j].re)), fma(c7, m6.v[j].im, fma(c5, m4.v[j].im, c3 * m2.v[j].im)));
      lv.v[j] = cx<T>(fma(c6, m6.v[j].re, fma(c4, m4.v[j].re, c2 * m2.v[j].re)), fma(c6, m6.v[j].im, fma(c4, m4.v[j].im, c2 * m2.v[j].im)));
    }
  }
  stage(sm, BA2, row, q, w);
  if (VJP) stage(sm, BM2, row, q, lw);
  __syncwarp();
  // U = A W;  Q = V - U, P = V + U;  Lu = A Lw + E W;  D1 = Lu + Lv, D2 = Lu - Lv
  Row<T, NT> qm, p, d1r, d2r;
  mul<T, NT, false>(sm, BA2, q, a, t);
#pragma unroll
  for (int j = 0; j < NT; ++j) { qm.v[j] = v.v[j] - t.v[j]; p.v[j] = v.v[j] + t.v[j]; }
  if (VJP) {
    mul<T, NT, false>(sm, BM2, q, a, t); mul<T, NT, true>(sm, BA2, q, e, t);
#pragma unroll
    for (int j = 0; j < NT; ++j) { d1r.v[j] = t.v[j] + lv.v[j]; d2r.v[j] = t.v[j] - lv.v[j]; }
  }
  __syncwarp();
  stage(sm, BA, row, q, p);
  // ---- Qi = Q^-1: unpivoted Gauss-Jordan, rows across the quad, the new pivot row broadcast by shuffle
#pragma unroll
  for (int k = 0; k < NT; ++k) {
    cx<T> prow[NT];
    const cx<T> pinv = crecip(quad_bcast(qm.v[k], qbase + k));
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const cx<T> x = quad_bcast(qm.v[j], qbase + k);
      prow[j] = (j == k) ? pinv : x * pinv;
    }
    const cx<T> h = qm.v[k];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      cx<T> base = (j == k) ? cx<T>(T(0), T(0)) : qm.v[j];
      cfma(base, -h, prow[j]);
      qm.v[j] = (row == k) ? prow[j] : base;
    }
  }
  __syncwarp();
  // R = Qi P;  G = D1 + D2 R;  L = Qi G
  Row<T, NT> r, l, g;
  mul<T, NT, false>(sm, BA, q, qm, r);
  constexpr int BR = VJP ? 1 : 1, BG = 2, BL = 3;
  stage(sm, BR, row, q, r);
  __syncwarp();
  if (VJP) {
    g = d1r;
    mul<T, NT, true>(sm, BR, q, d2r, g);
    stage(sm, BG, row, q, g);
    __syncwarp();
    mul<T, NT, false>(sm, BG, q, qm, l);
    stage(sm, BL, row, q, l);
    __syncwarp();
  }
  // ---- squarings (to the warp's largest s; a matrix past its own s keeps its result)
  for (int it = 0; it < smax; ++it) {
    Row<T, NT> rn, ln;
    mul<T, NT, false>(sm, BR, q, r, rn);
    if (VJP) { mul<T, NT, false>(sm, BL, q, r, ln); mul<T, NT, true>(sm, BR, q, l, ln); }
    __syncwarp();
    if (it < s) {
      r = rn; stage(sm, BR, row, q, r);
      if (VJP) { l = ln; stage(sm, BL, row, q, l); }
    }
    __syncwarp();
  }
  // ---- store Y = R and Abar = L^H (or L^T)
  const T nanv = bad ? (T)nan("") : T(0);
  if (live && row < n) {
#pragma unroll
    for (int j = 0; j < NT; ++j)
      if (j < n) {
        if (Yg) { cx<T> o = r.v[j]; o.re += nanv; o.im += nanv; Yg[(size_t)mat * nn + (size_t)row * n + j] = o; }
        if (VJP) {
          cx<T> o = l.v[j];
          if (conj_adjoint) o.im = -o.im;
          o.re += nanv; o.im += nanv;
          Abar[(size_t)mat * nn + (size_t)j * n + row] = o;
        }
      }
  }
}

template <typename T, int NT, bool VJP>
static int launch(const void* A, const void* Ybar, void* Y, void* Abar, int64_t batch, int n, int adjoint, cudaStream_t st) {
  constexpr int MPW = 32 / NT;
  const size_t smem = (size_t)kWarps * (VJP ? 4 : 2) * (NT * NT * MPW) * sizeof(cx<T>);
  const int64_t grid = (batch + kWarps * MPW - 1) / (kWarps * MPW);
  expm_tiny_kernel<T, NT, VJP><<<(unsigned)grid, kWarps * 32, smem, st>>>((const cx<T>*)A, (const cx<T>*)Ybar, (cx<T>*)Y, (cx<T>*)Abar,
                                                                   (long long)batch, n, adjoint == QG_ADJ_CONJ ? 1 : 0);
  QG_CHECK_LAUNCH();
  return QG_OK;
}

}  // namespace tiny

// Entry used by expm_small.cu: n <= 4 (both precisions), n <= 8 (complex64).
int expm_tiny(int dtype, const void* A, const void* Ybar, void* Y, void* Abar, int64_t batch, int n, int adjoint, bool vjp,
              cudaStream_t st) {
  if (dtype == QG_C128)
    return vjp ? tiny::launch<double, 4, true>(A, Ybar, Y, Abar, batch, n, adjoint, st)
               : tiny::launch<double, 4, false>(A, Ybar, Y, Abar, batch, n, adjoint, st);
  if (n > 4)      // complex64 up to n = 8: eight threads per matrix (a row of 8 complex64 is 16 registers)
    return vjp ? tiny::launch<float, 8, true>(A, Ybar, Y, Abar, batch, n, adjoint, st)
               : tiny::launch<float, 8, false>(A, Ybar, Y, Abar, batch, n, adjoint, st);
  return vjp ? tiny::launch<float, 4, true>(A, Ybar, Y, Abar, batch, n, adjoint, st)
             : tiny::launch<float, 4, false>(A, Ybar, Y, Abar, batch, n, adjoint, st);
}

}  // namespace qg

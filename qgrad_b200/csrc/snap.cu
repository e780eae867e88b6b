// Fused SNAP-gate cost (BASELINE config #3): examples/SNAP_gates.py:150-182 (apply_blocks) and :231-248 (cost),
//
//     x_T = prod_t [ D(-alpha_t) SNAP(theta_t) D(alpha_t) ] x_0 ,      cost = 1 - |<target | x_T>|^2 ,
//
// value AND gradient with respect to the complex alphas and the real thetas in ONE kernel, for batches of parameter sets
// that share the initial and the target state.  The reference rebuilds Displace (an eigh) on every call, forms six dense
// n x n operators and three diagonal ones per cost, and differentiates the chain with jax.grad; the composed path of this
// library (qgrad_qutip.apply_blocks: fused D.x kernels + torch glue) still wrote every intermediate state to HBM, 13 launches
// forward and as many backward.  Here a CTA owns 32 parameter sets (lane = set, the warps split the output indices), the state
// and its cotangent never leave shared memory, and nothing is taped: every factor is unitary on the truncated space
// (D(-alpha) = D(alpha)^H exactly, because the tridiagonal generator has a +/- symmetric spectrum), so the reverse sweep
// recovers each factor's INPUT by applying the inverse factor to its output -- the tape-free scheme the Givens kernels use.
//
//   D(alpha) x = conj(p) o V (w o V^T (p o x)),   p_k = s_k e^{-i phi k},  w_m = e^{i r lambda_m},  alpha = r e^{i phi}
//   D(-alpha): same w, p_k -> (-1)^k p_k.
//   reverse mode of y = D(beta) u with cotangent g:   ubar = D(beta)^H g,
//        rbar   = Re <g, conj(p) o V (i lambda o w o V^T (p o u))>
//        phibar = Re <g, i (k o y) - i conj(p) o V (w o V^T (k o p o u))>
//        betabar = (rbar b.re / r - phibar b.im / r^2,  rbar b.im / r + phibar b.re / r^2)      (0 at beta = 0, as Displace)
//   reverse mode of y = e^{i theta} o u:  thetabar_k = Im(conj(y_k)... ) = g.im y.re - g.re y.im,  ubar = e^{-i theta} o g.
// Gradients follow the library's convention (xbar = dL/dRe + i dL/dIm; jax.grad's is the conjugate, applied by the host).
// Work per set: 4 T matrix-vector products forward, 16 T in the reverse sweep (n^2 complex x real MACs each): FP64-FMA
// bound from n ~ 20 up; algorithmic HBM traffic is the parameters in and the cost + gradients out.
#include "common.cuh"

namespace qg {
namespace snap {

constexpr int kPitch = 33;      // vectors are stored [index][lane] with this pitch: lane-indexed accesses are conflict free

template <typename T> __device__ __forceinline__ cx<T> expi(T x) { T s, c; sincos_(x, &s, &c); return cx<T>(c, s); }
template <typename T> __device__ __forceinline__ cx<T> tscale(int j) { return (j & 1) ? cx<T>(T(0), T(1)) : cx<T>(T(1), T(0)); }

// out_j = sum_k M[k * np + j] in_k for this lane's sample, NV input vectors sharing the loads of M; the warp takes blocks of
// four output indices; fin(j, acc[NV]) consumes the results.
template <typename T, int NV, int kWarps, typename In, typename Fin>
__device__ __forceinline__ void matvec(const T* __restrict__ M, int n, int np, In in, Fin fin) {
  const int warp = threadIdx.x >> 5;
  for (int jb = warp; jb < np / 4; jb += kWarps) {
    cx<T> acc[NV][4];
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[v][u] = cx<T>(T(0), T(0));
#pragma unroll 4
    for (int k = 0; k < n; ++k) {
      T m[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) m[u] = M[k * np + jb * 4 + u];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const cx<T> x = in(v, k);
#pragma unroll
        for (int u = 0; u < 4; ++u) { acc[v][u].re = fma(m[u], x.re, acc[v][u].re); acc[v][u].im = fma(m[u], x.im, acc[v][u].im); }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = jb * 4 + u;
      if (j < n) { cx<T> r[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) r[v] = acc[v][u];
        fin(j, r); }
    }
  }
}

template <typename T>
struct Ctx {
  const T* sV; const T* sVt; const T* slam;       // V[k][m], V^T, lambda (padded to np)
  cx<T>* bP; cx<T>* bW;                           // p, w of the current block, [index][lane]
  cx<T>* bT; cx<T>* bU; cx<T>* bT2; cx<T>* bU2;   // scratch vectors
  int n, np, lane;
};

// dst <- D(+/- alpha) src   (sign = -1: p_k -> (-1)^k p_k).  dst may equal src.  Ends with a barrier.
template <typename T, int kWarps>
__device__ __forceinline__ void apply(const Ctx<T>& c, const cx<T>* src, cx<T>* dst, int sign) {
  const int n = c.n, np = c.np, lane = c.lane;
  auto pk = [&](int k) { cx<T> p = c.bP[k * kPitch + lane]; if (sign < 0 && (k & 1)) p = -p; return p; };
  for (int k = threadIdx.x >> 5; k < n; k += kWarps) c.bT[k * kPitch + lane] = pk(k) * src[k * kPitch + lane];
  __syncthreads();
  matvec<T, 1, kWarps>(c.sV, n, np, [&](int, int k) { return c.bT[k * kPitch + lane]; },
               [&](int m, cx<T>* r) { c.bU[m * kPitch + lane] = c.bW[m * kPitch + lane] * r[0]; });
  __syncthreads();
  matvec<T, 1, kWarps>(c.sVt, n, np, [&](int, int m) { return c.bU[m * kPitch + lane]; },
               [&](int j, cx<T>* r) { dst[j * kPitch + lane] = conj(pk(j)) * r[0]; });
  __syncthreads();
}

// Reverse mode of y = D(+/- alpha) u.  On entry bX holds y and bG its cotangent; on exit bX holds u and bG the cotangent of u.
// Returns this lane's (rbar, phibar) partial sums for the factor's OWN parameter beta = +/- alpha (warp-local: summed over
// the warp's output indices).
template <typename T, int kWarps>
__device__ __forceinline__ void unapply(const Ctx<T>& c, cx<T>* bX, cx<T>* bG, cx<T>* bY, int sign, T* rbar, T* phibar) {
  const int n = c.n, np = c.np, lane = c.lane;
  auto pk = [&](int k) { cx<T> p = c.bP[k * kPitch + lane]; if (sign < 0 && (k & 1)) p = -p; return p; };
  // keep y, recover u = D(beta)^H y = D(-beta) y
  for (int k = threadIdx.x >> 5; k < n; k += kWarps) bY[k * kPitch + lane] = bX[k * kPitch + lane];
  __syncthreads();
  apply<T, kWarps>(c, bY, bX, -sign);
  // t0 = p o u, t2 = k o t0;  u1 = w o V^T t0, u2 = w o V^T t2
  for (int k = threadIdx.x >> 5; k < n; k += kWarps) {
    const cx<T> t0 = pk(k) * bX[k * kPitch + lane];
    c.bT[k * kPitch + lane] = t0;
    c.bT2[k * kPitch + lane] = cx<T>((T)k * t0.re, (T)k * t0.im);
  }
  __syncthreads();
  matvec<T, 2, kWarps>(c.sV, n, np, [&](int v, int k) { return v ? c.bT2[k * kPitch + lane] : c.bT[k * kPitch + lane]; },
               [&](int m, cx<T>* r) {
                 const cx<T> w = c.bW[m * kPitch + lane];
                 c.bU[m * kPitch + lane] = w * r[0];
                 c.bU2[m * kPitch + lane] = w * r[1];
               });
  __syncthreads();
  // yr = conj(p) o V (i lambda o u1),  y2 = conj(p) o V u2;  rbar += Re <g, yr>,  phibar += Re <g, i (j y - y2)>
  T rb = T(0), pb = T(0);
  matvec<T, 2, kWarps>(c.sVt, n, np,
               [&](int v, int m) {
                 if (v) return c.bU2[m * kPitch + lane];
                 const cx<T> u = c.bU[m * kPitch + lane];
                 const T lm = c.slam[m];
                 return cx<T>(-lm * u.im, lm * u.re);
               },
               [&](int j, cx<T>* r) {
                 const cx<T> cp = conj(pk(j));
                 const cx<T> yr = cp * r[0], y2 = cp * r[1];
                 const cx<T> g = bG[j * kPitch + lane], y = bY[j * kPitch + lane];
                 rb += g.re * yr.re + g.im * yr.im;
                 const cx<T> z((T)j * y.re - y2.re, (T)j * y.im - y2.im);
                 pb += -(g.re * z.im - g.im * z.re);
               });
  *rbar = rb; *phibar = pb;
  __syncthreads();
  // cotangent of u: D(beta)^H g = D(-beta) g
  apply<T, kWarps>(c, bG, bG, -sign);
}

template <typename T, bool GRAD, int kWarps>
__global__ void __launch_bounds__(kWarps * 32) snap_cost_kernel(
    const T* __restrict__ V, const T* __restrict__ lam, const cx<T>* __restrict__ alphas, const T* __restrict__ thetas,
    const cx<T>* __restrict__ init, const cx<T>* __restrict__ target, T* __restrict__ cost, cx<T>* __restrict__ alphabar,
    T* __restrict__ thetabar, long long batch, int n, int np, int nblocks) {
  const int ntheta = n;          // full-length theta rows (pad_thetas, SNAP_gates.py:84-100): every factor stays unitary
  constexpr int kThreads = kWarps * 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sV = reinterpret_cast<T*>(smem_raw);
  T* sVt = sV + (size_t)n * np;
  T* slam = sVt + (size_t)n * np;
  T* sred = slam + np;                                    // [kWarps][32][2]
  cx<T>* sinit = reinterpret_cast<cx<T>*>(sred + kWarps * 64);
  cx<T>* starg = sinit + np;
  cx<T>* buf = starg + np;
  const size_t VB = (size_t)np * kPitch;
  Ctx<T> c;
  c.sV = sV; c.sVt = sVt; c.slam = slam; c.n = n; c.np = np;
  cx<T>* bX = buf; c.bP = buf + VB; c.bW = buf + 2 * VB; c.bT = buf + 3 * VB; c.bU = buf + 4 * VB;
  cx<T>* bG = buf + 5 * VB; cx<T>* bY = buf + 6 * VB; c.bT2 = buf + 7 * VB; c.bU2 = buf + 8 * VB;   // (the last four: GRAD only)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  c.lane = lane;
  const long long b0 = (long long)blockIdx.x * 32;
  const long long b = b0 + lane;
  const bool live = b < batch;
  for (int e = tid; e < n * np; e += kThreads) {
    const int k = e / np, m = e - k * np;
    sV[e] = m < n ? V[(size_t)k * n + m] : T(0);
    sVt[e] = m < n ? V[(size_t)m * n + k] : T(0);
  }
  for (int m = tid; m < np; m += kThreads) {
    slam[m] = m < n ? lam[m] : T(0);
    sinit[m] = m < n ? init[m] : cx<T>(T(0), T(0));
    starg[m] = m < n ? target[m] : cx<T>(T(0), T(0));
  }
  __syncthreads();
  for (int k = warp; k < n; k += kWarps) bX[k * kPitch + lane] = sinit[k];
  // p and w of block t (every warp recomputes its share; alpha = 0 -> r = 0, phi = 0 as Displace.__call__)
  auto set_block = [&](int t, cx<T>* a_out, T* r_out) {
    const cx<T> a = live ? alphas[(size_t)b * nblocks + t] : cx<T>(T(0), T(0));
    const T r = cabs_(a);
    const T phi = (r == T(0)) ? T(0) : atan2(a.im, a.re);
    __syncthreads();                                      // everyone is done with the previous block's p, w
    for (int k = warp; k < n; k += kWarps) {
      c.bP[k * kPitch + lane] = tscale<T>(k) * expi<T>(-phi * (T)k);
      c.bW[k * kPitch + lane] = expi<T>(r * slam[k]);
    }
    __syncthreads();
    *a_out = a; *r_out = r;
  };
  auto theta_at = [&](int t, int k) { return (live && k < ntheta) ? thetas[((size_t)b * nblocks + t) * ntheta + k] : T(0); };
  // ---- forward
  for (int t = 0; t < nblocks; ++t) {
    cx<T> a; T r;
    set_block(t, &a, &r);
    apply<T, kWarps>(c, bX, bX, +1);
    for (int k = warp; k < n; k += kWarps) bX[k * kPitch + lane] = expi<T>(theta_at(t, k)) * bX[k * kPitch + lane];
    __syncthreads();
    apply<T, kWarps>(c, bX, bX, -1);
  }
  // z = <target | x>, cost = 1 - |z|^2 ; every warp sums all indices for its lane (n is small)
  cx<T> z(T(0), T(0));
  for (int k = 0; k < n; ++k) cfma_conj(z, starg[k], bX[k * kPitch + lane]);
  if (warp == 0 && live) cost[b] = T(1) - (z.re * z.re + z.im * z.im);
  if (!GRAD) return;
  // ---- reverse sweep.  cotangent of x_T for cost = 1 - |z|^2:  -2 z target
  for (int k = warp; k < n; k += kWarps) bG[k * kPitch + lane] = (T(-2) * z) * starg[k];
  __syncthreads();
  for (int t = nblocks - 1; t >= 0; --t) {
    cx<T> a; T r;
    set_block(t, &a, &r);
    T rb_c, pb_c, rb_a, pb_a;
    unapply<T, kWarps>(c, bX, bG, bY, -1, &rb_c, &pb_c);             // factor D(-alpha_t)
    // SNAP: y = e^{i theta} o u
    for (int k = warp; k < n; k += kWarps) {
      const T th = theta_at(t, k);
      const cx<T> e = expi<T>(th);
      const cx<T> y = bX[k * kPitch + lane], g = bG[k * kPitch + lane];
      if (live && k < ntheta) thetabar[((size_t)b * nblocks + t) * ntheta + k] = g.im * y.re - g.re * y.im;
      bX[k * kPitch + lane] = conj(e) * y;
      bG[k * kPitch + lane] = conj(e) * g;
    }
    __syncthreads();
    unapply<T, kWarps>(c, bX, bG, bY, +1, &rb_a, &pb_a);             // factor D(alpha_t)
    // cross-warp sums of the four partials, in warp order (deterministic)
    T* s4 = sred;                                         // [4 values][kWarps][32]
    s4[(0 * kWarps + warp) * 32 + lane] = rb_a; s4[(1 * kWarps + warp) * 32 + lane] = pb_a;
    __syncthreads();
    T ra = T(0), pa = T(0);
    for (int w = 0; w < kWarps; ++w) { ra += s4[(0 * kWarps + w) * 32 + lane]; pa += s4[(1 * kWarps + w) * 32 + lane]; }
    __syncthreads();
    s4[(0 * kWarps + warp) * 32 + lane] = rb_c; s4[(1 * kWarps + warp) * 32 + lane] = pb_c;
    __syncthreads();
    T rc = T(0), pc = T(0);
    for (int w = 0; w < kWarps; ++w) { rc += s4[(0 * kWarps + w) * 32 + lane]; pc += s4[(1 * kWarps + w) * 32 + lane]; }
    if (warp == 0 && live) {
      // alphabar = dL/d alpha through D(alpha) minus the gradient w.r.t. beta = -alpha of D(beta)
      cx<T> ab(T(0), T(0));
      if (r > T(0)) {
        const T ir = T(1) / r, ir2 = ir * ir;
        ab = cx<T>(ra * a.re * ir - pa * a.im * ir2, ra * a.im * ir + pa * a.re * ir2);
        const cx<T> bt(-a.re, -a.im);
        const cx<T> bb(rc * bt.re * ir - pc * bt.im * ir2, rc * bt.im * ir + pc * bt.re * ir2);
        ab = ab - bb;
      }
      alphabar[(size_t)b * nblocks + t] = ab;
    }
  }
}

template <typename T, int kWarps>
static int run_w(const void* V, const void* lam, const void* alphas, const void* thetas, const void* init, const void* target,
                 void* cost, void* alphabar, void* thetabar, int64_t batch, int n, int nblocks, cudaStream_t st) {
  const int np = (n + 3) / 4 * 4;
  const bool grad = alphabar != nullptr;
  const size_t smem = (size_t)(2 * n * np + np + kWarps * 64) * sizeof(T) + (size_t)(2 * np) * sizeof(cx<T>) +
                      (size_t)(grad ? 9 : 5) * np * kPitch * sizeof(cx<T>) + 16;
  if (smem > 227 * 1024) return QG_ERR_DIM;
  auto kg = snap_cost_kernel<T, true, kWarps>; auto kv = snap_cost_kernel<T, false, kWarps>;
  if (smem > 48 * 1024) QG_RETURN_IF_CUDA(cudaFuncSetAttribute(grad ? kg : kv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long grid = (batch + 31) / 32;
  if (grid > 0x7fffffffLL) return QG_ERR_ARG;
  if (grad) kg<<<(unsigned)grid, kWarps * 32, smem, st>>>((const T*)V, (const T*)lam, (const cx<T>*)alphas, (const T*)thetas, (const cx<T>*)init,
                                                         (const cx<T>*)target, (T*)cost, (cx<T>*)alphabar, (T*)thetabar, batch, n, np, nblocks);
  else kv<<<(unsigned)grid, kWarps * 32, smem, st>>>((const T*)V, (const T*)lam, (const cx<T>*)alphas, (const T*)thetas, (const cx<T>*)init,
                                                    (const cx<T>*)target, (T*)cost, nullptr, nullptr, batch, n, np, nblocks);
  QG_CHECK_LAUNCH();
  return QG_OK;
}
// warps per CTA: the output indices come in blocks of four, one block per warp and round -- four warps up to n = 16
// (574 vs 702 us at n = 10 with eight), eight above (5958 -> 4733 us at n = 40)
template <typename T>
static int run(const void* V, const void* lam, const void* alphas, const void* thetas, const void* init, const void* target,
               void* cost, void* alphabar, void* thetabar, int64_t batch, int n, int nblocks, cudaStream_t st) {
  if (n <= 16) return run_w<T, 4>(V, lam, alphas, thetas, init, target, cost, alphabar, thetabar, batch, n, nblocks, st);
  // one round of output blocks per matrix-vector product: with 9 or 10 blocks (n = 33 ... 40) eight warps ran two rounds, the
  // second with two warps busy
  const int blocks = (n + 3) / 4;
  if (blocks <= 8) return run_w<T, 8>(V, lam, alphas, thetas, init, target, cost, alphabar, thetabar, batch, n, nblocks, st);
  if (blocks <= 10) return run_w<T, 10>(V, lam, alphas, thetas, init, target, cost, alphabar, thetabar, batch, n, nblocks, st);
  if (blocks <= 12) return run_w<T, 12>(V, lam, alphas, thetas, init, target, cost, alphabar, thetabar, batch, n, nblocks, st);
  return run_w<T, 16>(V, lam, alphas, thetas, init, target, cost, alphabar, thetabar, batch, n, nblocks, st);
}

}  // namespace snap

int snap_cost(int dtype, const void* V, const void* lam, const void* alphas, const void* thetas, const void* init, const void* target,
              void* cost, void* alphabar, void* thetabar, int64_t batch, int n, int nblocks, cudaStream_t st) {
  return dtype == QG_C128 ? snap::run<double>(V, lam, alphas, thetas, init, target, cost, alphabar, thetabar, batch, n, nblocks, st)
                          : snap::run<float>(V, lam, alphas, thetas, init, target, cost, alphabar, thetabar, batch, n, nblocks, st);
}

}  // namespace qg

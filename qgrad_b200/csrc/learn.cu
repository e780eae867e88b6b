// K7 -- the pieces of one Hamiltonian-learning step around the expm forward / backward passes.
// Workload definition: examples/Unitary-Learning-qgrad.py:123-151 and
// examples/Unitary-Learning-no-qgrad.py:112-191 (cost 1 - mean_k |Re <out_k|U|in_k>|), with
// U_s = exp(-i t H_s(theta)), H_s = sum_p ctrl[s,p] theta_p P_p over Pauli strings.
// The ket products (Phi = U Psi_in, Ubar = G Psi_in^H) run on the GEMM kernel of gemm.cuh.
#include "common.cuh"
#include <cooperative_groups.h>
#include <cstdint>

namespace qg {
namespace learn {

// <r|P|c> = i^popcount(x&z) (-1)^popcount(c&z) [r == c^x]
template <typename T> __device__ __forceinline__ cx<T> pauli_phase(int x, int z, int c) {
  const int q = (__popc(x & z) + 2 * __popc(c & z)) & 3;
  return q == 0 ? cx<T>(T(1), T(0)) : q == 1 ? cx<T>(T(0), T(1)) : q == 2 ? cx<T>(T(-1), T(0)) : cx<T>(T(0), T(-1));
}

// A block owns kHamRows consecutive rows (of the n_set * n rows of all settings): it zeroes them with 16-byte
// stores, then one thread per (row, term p) writes entry (r, r ^ x_p) for the FIRST term carrying that flip mask,
// summing the later terms with the same mask in term order (X/Y on the same qubits, all Z-strings on the diagonal):
// A_s[r][r^x] = sum_{p: x_p = x} (-i t) ctrl[s,p] theta_p phase_p.  One pass over A, no read-modify-write, no
// atomics, deterministic; `zero` = 0 when the caller cleared A (unaligned A).
constexpr int kHamRows = 4;
template <typename T>
__global__ void __launch_bounds__(256)
pauli_hamiltonian_kernel(const T* __restrict__ theta, const T* __restrict__ ctrl, const int* __restrict__ xm,
                         const int* __restrict__ zm, int n_terms, int n_set, int n, T t, cx<T>* __restrict__ A, int zero) {
  const long long rows = (long long)n_set * n;
  const long long g0 = (long long)blockIdx.x * kHamRows;
  const int nr = (int)(rows - g0 < kHamRows ? rows - g0 : kHamRows);
  if (zero) {
    uint4* z = reinterpret_cast<uint4*>(A + (size_t)g0 * n);
    const int units = (int)((size_t)nr * n * sizeof(cx<T>) / 16);
    for (int u = threadIdx.x; u < units; u += blockDim.x) z[u] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
  }
  for (int w = threadIdx.x; w < nr * n_terms; w += blockDim.x) {
    const int lr = w / n_terms, p = w - lr * n_terms;
    const long long gr = g0 + lr;
    const int s = (int)(gr / n), r = (int)(gr % n);
    const int x = xm[p];
    bool owner = true;
    for (int q = 0; q < p; ++q) owner &= (xm[q] != x);             // an earlier term owns this entry
    if (!owner) continue;
    const int c = r ^ x;
    cx<T> v(T(0), T(0));
    for (int q = p; q < n_terms; ++q) {
      if (xm[q] != x) continue;
      const T wgt = t * ctrl[(size_t)s * n_terms + q] * theta[q];
      const cx<T> ph = pauli_phase<T>(x, zm[q], c);
      v.re += wgt * ph.im; v.im -= wgt * ph.re;                    // (-i) * ph * w = (ph.im, -ph.re) * w
    }
    A[(size_t)gr * n + c] = v;
  }
}

// thetabar[p] (+)= sum_s ctrl[s,p] sum_r Re( conj(Abar_s[r][c]) * (-i t) phase_p(c) ),  c = r ^ x_p
template <typename T>
__global__ void pauli_project_kernel(const cx<T>* __restrict__ Abar, const T* __restrict__ ctrl, const int* __restrict__ xm,
                                     const int* __restrict__ zm, int n_terms, int n_set, int n, T t,
                                     T* __restrict__ thetabar, int accumulate) {
  __shared__ T red[32];
  const int p = blockIdx.x;
  const int x = xm[p], z = zm[p];
  T acc = T(0);
#pragma unroll 8
  for (long long g = threadIdx.x; g < (long long)n_set * n; g += blockDim.x) {
    const int s = (int)(g / n), r = (int)(g % n), c = r ^ x;
    const cx<T> a = Abar[(size_t)s * n * n + (size_t)r * n + c];
    const cx<T> ph = pauli_phase<T>(x, z, c);
    const cx<T> d(ph.im, -ph.re);                               // -i * phase
    acc += ctrl[(size_t)s * n_terms + p] * (a.re * d.re + a.im * d.im);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    T tot = T(0);
    for (int w = 0; w < (int)(blockDim.x + 31) / 32; ++w) tot += red[w];
    tot *= t;
    thetabar[p] = accumulate ? thetabar[p] + tot : tot;
  }
}

// block (32 kets, 8 row lanes): o_k = sum_i conj(out_ik) phi_ik ; loss += sum_k |Re o_k| ;
// G_ik = -sign(Re o_k) inv_m out_ik.    grid (m_local/32, batch)
template <typename T>
__global__ void overlap_loss_kernel(const cx<T>* __restrict__ phi, const cx<T>* __restrict__ out, cx<T>* __restrict__ G,
                                    T* __restrict__ loss, int n, int m_local, T inv_m) {
  __shared__ T part[8][32];
  __shared__ T gk[32];
  const int kx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int k = blockIdx.x * 32 + kx;
  const size_t base = (size_t)blockIdx.y * n * m_local;
  T re = T(0);
  if (k < m_local)
    for (int i = ry; i < n; i += 8) {
      const cx<T> o = out[base + (size_t)i * m_local + k], f = phi[base + (size_t)i * m_local + k];
      re += o.re * f.re + o.im * f.im;                           // Re(conj(o) f)
    }
  part[ry][kx] = re;
  __syncthreads();
  if (ry == 0) {
    T tot = T(0);
#pragma unroll
    for (int q = 0; q < 8; ++q) tot += part[q][kx];
    const T sg = tot > T(0) ? T(1) : (tot < T(0) ? T(-1) : T(0));
    gk[kx] = (k < m_local) ? -sg * inv_m : T(0);
    T a = (k < m_local) ? fabs(tot) : T(0);
    a = warp_sum(a);
    if (kx == 0) atomicAdd(loss, a);
  }
  __syncthreads();
  if (k < m_local) {
    const T g = gk[kx];
    for (int i = ry; i < n; i += 8) {
      const cx<T> o = out[base + (size_t)i * m_local + k];
      G[base + (size_t)i * m_local + k] = cx<T>(g * o.re, g * o.im);
    }
  }
}

// Same result through a thread-block cluster: the kOvSplit CTAs of a cluster share one block of 32 kets, CTA c owns
// the rows [c * ceil(n / kOvSplit), ...), keeps its `out` entries in registers across the reduction (no second read),
// and the per-ket totals meet through distributed shared memory.  8x the CTAs of the kernel above (the 8-rank local
// step has batch = 1: 16 CTAs there, 128 here) and 3 instead of 4 passes over HBM.
constexpr int kOvSplit = 8, kOvRows = 128;                         // rows per CTA; n <= kOvSplit * kOvRows
template <typename T, int RL, int MINB>                              // RL row lanes: 32 * RL threads, kOvRows / RL rows per thread
__global__ void __cluster_dims__(kOvSplit, 1, 1) __launch_bounds__(32 * RL, MINB)
overlap_loss_cluster_kernel(const cx<T>* __restrict__ phi, const cx<T>* __restrict__ out, cx<T>* __restrict__ G,
                            T* __restrict__ loss, int n, int m_local, T inv_m) {
  namespace cg = cooperative_groups;
  constexpr int Q = kOvRows / RL;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ T part[RL][32];
  __shared__ T mine[32];
  __shared__ T gk[32];
  const int kx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int k = blockIdx.y * 32 + kx;
  const size_t base = (size_t)blockIdx.z * n * m_local;
  const int crank = (int)cluster.block_rank();
  const int rows_per = (n + kOvSplit - 1) / kOvSplit;
  const int r0 = crank * rows_per, r1 = r0 + rows_per < n ? r0 + rows_per : n;
  cx<T> o[Q];
  T re = T(0);
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int i = r0 + ry + RL * q;
    o[q] = cx<T>(T(0), T(0));
    if (k < m_local && i < r1) {
      o[q] = out[base + (size_t)i * m_local + k];
      const cx<T> f = phi[base + (size_t)i * m_local + k];
      re += o[q].re * f.re + o[q].im * f.im;                      // Re(conj(o) f)
    }
  }
  part[ry][kx] = re;
  __syncthreads();
  if (ry == 0) {
    T tot = T(0);
#pragma unroll
    for (int q = 0; q < RL; ++q) tot += part[q][kx];
    mine[kx] = tot;
  }
  cluster.sync();
  if (ry == 0) {
    T tot = T(0);
    for (int c = 0; c < kOvSplit; ++c) tot += *cluster.map_shared_rank(&mine[kx], c);   // same order in every CTA
    const T sg = tot > T(0) ? T(1) : (tot < T(0) ? T(-1) : T(0));
    gk[kx] = (k < m_local) ? -sg * inv_m : T(0);
    if (crank == 0) {
      T a = (k < m_local) ? fabs(tot) : T(0);
      a = warp_sum(a);
      if (kx == 0) atomicAdd(loss, a);
    }
  }
  cluster.sync();                                                  // gk visible; nobody leaves while `mine` is still being read
  const T g = gk[kx];
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int i = r0 + ry + RL * q;
    if (k < m_local && i < r1) G[base + (size_t)i * m_local + k] = cx<T>(g * o[q].re, g * o[q].im);
  }
}

// dst[b] (cols x rows) = src[b]^H, src (rows x cols): 32x32 tiles through shared memory, both sides coalesced.
// grid (ceil(cols/32), ceil(rows/32), batch), block (32, 8)
template <typename T>
__global__ void conj_transpose_kernel(const cx<T>* __restrict__ src, cx<T>* __restrict__ dst, int rows, int cols) {
  __shared__ cx<T> tile[32][33];
  const size_t base = (size_t)blockIdx.z * rows * cols;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[i][threadIdx.x] = src[base + (size_t)r * cols + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + threadIdx.x;                       // dst row c, dst column r
    if (r < rows && c < cols) { const cx<T> v = tile[threadIdx.x][i]; dst[base + (size_t)c * rows + r] = cx<T>(v.re, -v.im); }
  }
}

template <typename T>
__global__ void adam_kernel(T* __restrict__ x, T* __restrict__ m, T* __restrict__ v, const T* __restrict__ g,
                            long long count, T lr, T b1, T b2, T eps, T c1, T c2) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const T gi = g[i];
  const T mi = (T(1) - b1) * gi + b1 * m[i];
  const T vi = (T(1) - b2) * gi * gi + b2 * v[i];
  m[i] = mi; v[i] = vi;
  x[i] = x[i] - lr * (mi / c1) / (sqrt(vi / c2) + eps);
}

// Graph-capturable tail of a training step: Adam with the step index read from (and advanced in) DEVICE memory, so
// the same captured launch serves every step, plus the step's loss 1 - loss_partial / m_total written next to it.
// One block (the parameter vectors of this path are tens to thousands of reals).
template <typename T>
__global__ void adam_device_kernel(T* __restrict__ x, T* __restrict__ m, T* __restrict__ v, const T* __restrict__ g,
                                   long long count, T lr, T b1, T b2, T eps, long long* __restrict__ step,
                                   const T* __restrict__ loss_partial, T inv_m_total, T* __restrict__ loss_out) {
  const long long i = *step;
  const T c1 = T(1) - (T)pow((double)b1, (double)(i + 1)), c2 = T(1) - (T)pow((double)b2, (double)(i + 1));
  for (long long k = threadIdx.x; k < count; k += blockDim.x) {
    const T gi = g[k];
    const T mi = (T(1) - b1) * gi + b1 * m[k];
    const T vi = (T(1) - b2) * gi * gi + b2 * v[k];
    m[k] = mi; v[k] = vi;
    x[k] = x[k] - lr * (mi / c1) / (sqrt(vi / c2) + eps);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    *step = i + 1;
    if (loss_out) *loss_out = T(1) - *loss_partial * inv_m_total;
  }
}

template <typename T>
static int hamiltonian(const void* theta, const void* ctrl, const int* xm, const int* zm, int n_terms, int n_set, int nq,
                       double t, void* A, cudaStream_t st) {
  const int n = 1 << nq;
  const bool aligned = ((uintptr_t)A & 15) == 0;                   // n >= 2: every row block is a multiple of 16 bytes
  if (!aligned) QG_RETURN_IF_CUDA(cudaMemsetAsync(A, 0, (size_t)n_set * n * n * sizeof(cx<T>), st));
  const long long rows = (long long)n_set * n;
  pauli_hamiltonian_kernel<T><<<(unsigned)((rows + kHamRows - 1) / kHamRows), 256, 0, st>>>((const T*)theta, (const T*)ctrl, xm, zm, n_terms, n_set, n, (T)t, (cx<T>*)A, aligned ? 1 : 0);
  QG_CHECK_LAUNCH();
  return QG_OK;
}
template <typename T>
static int project(const void* Abar, const void* ctrl, const int* xm, const int* zm, int n_terms, int n_set, int nq,
                   double t, void* tb, int accumulate, cudaStream_t st) {
  const int n = 1 << nq;
  pauli_project_kernel<T><<<n_terms, 1024, 0, st>>>((const cx<T>*)Abar, (const T*)ctrl, xm, zm, n_terms, n_set, n, (T)t, (T*)tb, accumulate);
  QG_CHECK_LAUNCH();
  return QG_OK;
}
template <typename T>
static int overlap(const void* phi, const void* out, void* G, void* loss, int64_t batch, int n, int m_local, double inv_m,
                   cudaStream_t st) {
  if (n <= kOvSplit * kOvRows) {
    // 16 row lanes (512 threads, 8 rows per thread) measured best: 8 lanes need 128 registers in complex128 (2 CTAs
    // per SM, 33 resident clusters), 32 lanes gain nothing more.  1024 x (8 x 512) complex128: 94 -> 57 us; one
    // setting (the 8-rank shard): 67 -> 14 us.
    dim3 cgrid(kOvSplit, (m_local + 31) / 32, (unsigned)batch);
    overlap_loss_cluster_kernel<T, 16, (sizeof(T) == 8 ? 2 : 4)><<<cgrid, 512, 0, st>>>((const cx<T>*)phi, (const cx<T>*)out, (cx<T>*)G, (T*)loss, n, m_local, (T)inv_m);
    QG_CHECK_LAUNCH();
    return QG_OK;
  }
  dim3 grid((m_local + 31) / 32, (unsigned)batch);
  overlap_loss_kernel<T><<<grid, 256, 0, st>>>((const cx<T>*)phi, (const cx<T>*)out, (cx<T>*)G, (T*)loss, n, m_local, (T)inv_m);
  QG_CHECK_LAUNCH();
  return QG_OK;
}
template <typename T>
static int adam(void* x, void* m, void* v, const void* g, int64_t count, double lr, double b1, double b2, double eps,
                int64_t i, cudaStream_t st) {
  const double c1 = 1.0 - pow(b1, (double)(i + 1)), c2 = 1.0 - pow(b2, (double)(i + 1));
  adam_kernel<T><<<(unsigned)((count + 255) / 256), 256, 0, st>>>((T*)x, (T*)m, (T*)v, (const T*)g, count, (T)lr, (T)b1, (T)b2, (T)eps, (T)c1, (T)c2);
  QG_CHECK_LAUNCH();
  return QG_OK;
}

template <typename T>
static int adam_device(void* x, void* m, void* v, const void* g, int64_t count, double lr, double b1, double b2, double eps,
                       int64_t* step, const void* loss_partial, double inv_m_total, void* loss_out, cudaStream_t st) {
  adam_device_kernel<T><<<1, 256, 0, st>>>((T*)x, (T*)m, (T*)v, (const T*)g, count, (T)lr, (T)b1, (T)b2, (T)eps, (long long*)step,
                                           (const T*)loss_partial, (T)inv_m_total, (T*)loss_out);
  QG_CHECK_LAUNCH();
  return QG_OK;
}

template <typename T>
static int conj_transpose(const void* src, void* dst, int64_t batch, int rows, int cols, cudaStream_t st) {
  if (batch > 65535) return QG_ERR_ARG;
  dim3 grid((cols + 31) / 32, (rows + 31) / 32, (unsigned)batch);
  conj_transpose_kernel<T><<<grid, dim3(32, 8), 0, st>>>((const cx<T>*)src, (cx<T>*)dst, rows, cols);
  QG_CHECK_LAUNCH();
  return QG_OK;
}

}  // namespace learn

int learn_hamiltonian(int dtype, const void* theta, const void* ctrl, const int* xm, const int* zm, int n_terms, int n_set,
                      int nq, double t, void* A, cudaStream_t st) {
  return dtype == QG_C128 ? learn::hamiltonian<double>(theta, ctrl, xm, zm, n_terms, n_set, nq, t, A, st)
                          : learn::hamiltonian<float>(theta, ctrl, xm, zm, n_terms, n_set, nq, t, A, st);
}
int learn_project(int dtype, const void* Abar, const void* ctrl, const int* xm, const int* zm, int n_terms, int n_set, int nq,
                  double t, void* tb, int accumulate, cudaStream_t st) {
  return dtype == QG_C128 ? learn::project<double>(Abar, ctrl, xm, zm, n_terms, n_set, nq, t, tb, accumulate, st)
                          : learn::project<float>(Abar, ctrl, xm, zm, n_terms, n_set, nq, t, tb, accumulate, st);
}
int learn_overlap(int dtype, const void* phi, const void* out, void* G, void* loss, int64_t batch, int n, int m_local,
                  double inv_m, cudaStream_t st) {
  return dtype == QG_C128 ? learn::overlap<double>(phi, out, G, loss, batch, n, m_local, inv_m, st)
                          : learn::overlap<float>(phi, out, G, loss, batch, n, m_local, inv_m, st);
}
int learn_conj_transpose(int dtype, const void* src, void* dst, int64_t batch, int rows, int cols, cudaStream_t st) {
  return dtype == QG_C128 ? learn::conj_transpose<double>(src, dst, batch, rows, cols, st)
                          : learn::conj_transpose<float>(src, dst, batch, rows, cols, st);
}
int learn_adam(int dtype, void* x, void* m, void* v, const void* g, int64_t count, double lr, double b1, double b2,
               double eps, int64_t i, cudaStream_t st) {
  return dtype == QG_C128 ? learn::adam<double>(x, m, v, g, count, lr, b1, b2, eps, i, st)
                          : learn::adam<float>(x, m, v, g, count, lr, b1, b2, eps, i, st);
}
int learn_adam_device(int dtype, void* x, void* m, void* v, const void* g, int64_t count, double lr, double b1, double b2,
                      double eps, int64_t* step, const void* loss_partial, double inv_m_total, void* loss_out, cudaStream_t st) {
  return dtype == QG_C128 ? learn::adam_device<double>(x, m, v, g, count, lr, b1, b2, eps, step, loss_partial, inv_m_total, loss_out, st)
                          : learn::adam_device<float>(x, m, v, g, count, lr, b1, b2, eps, step, loss_partial, inv_m_total, loss_out, st);
}

}  // namespace qg

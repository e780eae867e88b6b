// Device-side dataset / state generators: Haar-random kets, GUE Hamiltonians, uniform angles
// for rand_unitary, and the reference's own rand_ket structure -- replacing the host-side QuTiP / tenpy / jax.random
// draws of qgrad/qgrad_qutip.py:558-620 and examples/Unitary-Learning-qgrad.py:63,88,
// examples/Unitary-Learning-no-qgrad.py:106-107, so that a benchmark or a training script never stages its synthetic
// data through the host.
//
// Counter-based Philox4x32-10 (Salmon et al., SC'11): the value at (seed, stream, index) is a pure function of its
// coordinates -- no generator state, any element can be (re)computed by any thread, results do not depend on the
// launch geometry.  JAX's threefry stream is not reproduced (it cannot be checked here): the reference's
// tests only pin PROPERTIES of these draws (unit norm, unitarity, Hermiticity, trace one), and so do ours, plus the
// moments of the distributions.
//   uniform  : u = (x + 0.5) 2^-32                     in (0, 1), one 32-bit word (two words -> 53 bits for float64)
//   normal   : Box-Muller on two uniforms -> two independent N(0, 1)
#include "common.cuh"

namespace qg {
namespace rnd {

struct u4 { unsigned x, y, z, w; };

__host__ __device__ __forceinline__ u4 philox4x32_10(u4 ctr, unsigned k0, unsigned k1) {
  constexpr unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned long long p0 = (unsigned long long)M0 * ctr.x, p1 = (unsigned long long)M1 * ctr.z;
    const unsigned hi0 = (unsigned)(p0 >> 32), lo0 = (unsigned)p0, hi1 = (unsigned)(p1 >> 32), lo1 = (unsigned)p1;
    ctr = u4{hi1 ^ ctr.y ^ k0, lo1, hi0 ^ ctr.w ^ k1, lo0};
    k0 += W0; k1 += W1;
  }
  return ctr;
}

// four words at (seed, stream, index)
__device__ __forceinline__ u4 draw(unsigned long long seed, unsigned stream, unsigned long long index) {
  return philox4x32_10(u4{(unsigned)index, (unsigned)(index >> 32), stream, 0x51677261u /* "qgra" */},
                       (unsigned)seed, (unsigned)(seed >> 32));
}

template <typename T> __device__ __forceinline__ T unit_open(unsigned a, unsigned b);
template <> __device__ __forceinline__ float unit_open<float>(unsigned a, unsigned) {
  return ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f);                       // 24 bits: exactly representable, never 0 or 1
}
template <> __device__ __forceinline__ double unit_open<double>(unsigned a, unsigned b) {
  const unsigned long long v = ((unsigned long long)a << 21) ^ (unsigned long long)(b >> 11);   // 53 bits
  return ((double)v + 0.5) * (1.0 / 9007199254740992.0);
}

// two independent standard normals from four words
template <typename T> __device__ __forceinline__ cx<T> normal_pair(u4 w) {
  const T u1 = unit_open<T>(w.x, w.y), u2 = unit_open<T>(w.z, w.w);
  const T r = sqrt(T(-2) * log(u1));
  T s, c;
  sincos_(T(6.283185307179586476925286766559) * u2, &s, &c);
  return cx<T>(r * c, r * s);
}

enum { STREAM_KET = 1, STREAM_HERM = 2, STREAM_UNIFORM = 3 };

// ---- kets: one CTA per ket.  kind 0: Haar (normalised complex Gaussian); kind 1: the reference's rand_ket
// (qgrad_qutip.py:570-573: uniform(key) + 1j * uniform(key) with the SAME key, i.e. re == im, normalised).
// Two passes over the counter space (norm, then scaled store) instead of two passes over memory.
template <typename T>
__global__ void __launch_bounds__(256) rand_ket_kernel(cx<T>* __restrict__ out, int n, unsigned long long seed,
                                                       unsigned long long first_ket, int kind) {
  __shared__ T red[8];
  const unsigned long long ket = first_ket + blockIdx.x;
  auto elem = [&](int i) {
    const u4 w = draw(seed, STREAM_KET, ket * (unsigned long long)n + i);
    if (kind == 1) { const T u = unit_open<T>(w.x, w.y); return cx<T>(u, u); }
    return normal_pair<T>(w);
  };
  T nn = T(0);
  for (int i = threadIdx.x; i < n; i += blockDim.x) { const cx<T> v = elem(i); nn = fma(v.re, v.re, fma(v.im, v.im, nn)); }
  nn = warp_sum(nn);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = nn;
  __syncthreads();
  T tot = T(0);
  for (int w = 0; w < 8; ++w) tot += red[w];                                     // fixed order: reproducible
  const T inv = T(1) / sqrt(tot);
  cx<T>* o = out + (size_t)blockIdx.x * n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { const cx<T> v = elem(i); o[i] = cx<T>(v.re * inv, v.im * inv); }
}

// ---- Hermitian matrices H = scale (X + X^H) / 2, X iid complex standard normal (GUE up to the scale; SURVEY 8d:
// scale = 1 / sqrt(n) gives the named H / sqrt(n)).  Thread per (r, c): both X_rc and X_cr are functions of their
// coordinates, so every entry is written once, in memory order, and H is Hermitian bit for bit.
template <typename T>
__global__ void rand_hermitian_kernel(cx<T>* __restrict__ H, int n, unsigned long long seed, unsigned long long first, T scale) {
  const size_t nn = (size_t)n * n;
  const unsigned long long mat = first + blockIdx.y;
  cx<T>* h = H + (size_t)blockIdx.y * nn;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < nn; e += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / n), c = (int)(e % n);
    const cx<T> a = normal_pair<T>(draw(seed, STREAM_HERM, mat * nn + (size_t)r * n + c));
    const cx<T> b = normal_pair<T>(draw(seed, STREAM_HERM, mat * nn + (size_t)c * n + r));
    const T half = T(0.5) * scale;
    h[e] = r == c ? cx<T>(scale * a.re, T(0)) : cx<T>(half * (a.re + b.re), half * (a.im - b.im));
  }
}

// ---- uniform reals in [lo, hi): the N^2 angles of rand_unitary (qgrad_qutip.py:612-620), parameter sets for sweeps
template <typename T>
__global__ void rand_uniform_kernel(T* __restrict__ out, long long count, unsigned long long seed, T lo, T hi) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    const u4 w = draw(seed, STREAM_UNIFORM, (unsigned long long)i);
    out[i] = fma(unit_open<T>(w.x, w.y), hi - lo, lo);
  }
}

template <typename T>
static int run(int what, void* out, int64_t batch, int n, uint64_t seed, uint64_t first, int kind, double a, double b, cudaStream_t st) {
  switch (what) {
    case 0:
      if (batch > 0x7fffffffLL) return QG_ERR_ARG;
      rand_ket_kernel<T><<<(unsigned)batch, 256, 0, st>>>((cx<T>*)out, n, seed, first, kind);
      break;
    case 1: {
      if (batch > 65535) return QG_ERR_ARG;
      const size_t nn = (size_t)n * n;
      const unsigned gx = (unsigned)((nn + 255) / 256 < 1024 ? (nn + 255) / 256 : 1024);
      rand_hermitian_kernel<T><<<dim3(gx, (unsigned)batch), 256, 0, st>>>((cx<T>*)out, n, seed, first, (T)a);
      break;
    }
    default: {
      const long long blocks = (batch + 255) / 256;
      rand_uniform_kernel<T><<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, st>>>((T*)out, batch, seed, (T)a, (T)b);
    }
  }
  QG_CHECK_LAUNCH();
  return QG_OK;
}

}  // namespace rnd

int rand_op(int dtype, int what, void* out, int64_t batch, int n, uint64_t seed, uint64_t first, int kind, double a, double b,
            cudaStream_t st) {
  return dtype == QG_C128 ? rnd::run<double>(what, out, batch, n, seed, first, kind, a, b, st)
                          : rnd::run<float>(what, out, batch, n, seed, first, kind, a, b, st);
}

}  // namespace qg

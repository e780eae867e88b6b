// Batched complex GEMM with fused multi-term accumulation and linear-combination epilogues:
//
//   acc = A0.B0 (+ A1.B1)
//   C1  = al1*acc + sg1*lin,   C2 = al2*acc + sg2*lin        (C2 optional)
//   lin = x0*X0 + x1*X1 + x2*X2 + xI*I
//
// This one kernel carries every product of the large-dimension Pade / Frechet recurrences
// (A^2, A^4, A^6 with W as second output, U with Q = V-U and P = V+U as the two outputs, the
// dual products A.E + E.A, the squarings R.R and R.L + L.R), the rank-64 updates of the blocked
// Gauss-Jordan inverse, and the ket products of the learning step.  Tasks whose result is known to be Hermitian,
// skew-Hermitian or a (V - U, V + U) pair (Task::herm) compute the upper triangle of tiles and write the mirror.
//
// complex128: 64x64 CTA tile, 4 warps, 32x32 warp tile of FP64 tensor MMAs (mma.sync m8n8k4 =
// SASS DMMA.8x8x4; four real MMAs per complex 8x8x4 step), BK = 16, 3-stage cp.async pipeline,
// XOR-swizzled shared tiles so both fragment loads are conflict-free 16-byte LDS.
// complex64: same tiling, FP32 FMAs, 8x4 complex micro-tile per thread.
#pragma once
#include <stdlib.h>
#include <type_traits>
#include "common.cuh"

namespace qg {
namespace gemm {

constexpr int BM = 64, BN = 64, BK = 16, STAGES = 3, NTHREADS = 128;

template <typename T>
struct Task {
  const cx<T>* A0; const cx<T>* B0; const cx<T>* A1; const cx<T>* B1;
  cx<T>* C1; cx<T>* C2;
  const cx<T>* X0; const cx<T>* X1; const cx<T>* X2;
  T al1, sg1, al2, sg2, x0, x1, x2, xI;
  long long sA0, sB0, sA1, sB1, sC1, sC2, sX0, sX1, sX2;   // batch strides in elements
  int lda0, ldb0, lda1, ldb1, ldc1, ldc2, ldx0, ldx1, ldx2; // row pitches in elements
  int M, N, K;               // multiples of 64 / 64 / 16
  int skip_tm, skip_tn;      // tile row / col to leave untouched (-1: none)
  int skip_r0, skip_c0;      // tensor-core kernel only: first row / column of a 128-wide band to leave untouched (-1: none)
  int pred_it;               // >= 0: run for matrix b only if pred_it < s[b]
  int herm;                  // structure hint, used when sym[b] != 0 (square tasks only): 1 = C1 (and C2) are Hermitian;
                             // 2 = C1 = V - U, C2 = V + U with V Hermitian and U skew-Hermitian (sym 1) or Hermitian (sym 2);
                             // 3 = C1 (and C2) are skew-Hermitian (honoured for sym 1 only)
};

template <typename T>
inline Task<T> make_task() {
  Task<T> t;
  t.A0 = t.B0 = t.A1 = t.B1 = nullptr; t.C1 = t.C2 = nullptr; t.X0 = t.X1 = t.X2 = nullptr;
  t.al1 = 1; t.sg1 = 1; t.al2 = 1; t.sg2 = 1; t.x0 = t.x1 = t.x2 = t.xI = 0;
  t.sA0 = t.sB0 = t.sA1 = t.sB1 = t.sC1 = t.sC2 = t.sX0 = t.sX1 = t.sX2 = 0;
  t.lda0 = t.ldb0 = t.lda1 = t.ldb1 = t.ldc1 = t.ldc2 = t.ldx0 = t.ldx1 = t.ldx2 = 0;
  t.M = t.N = t.K = 0; t.skip_tm = t.skip_tn = -1; t.skip_r0 = t.skip_c0 = -1; t.pred_it = -1; t.herm = 0;
  return t;
}

template <typename T>
struct Launch {
  Task<T> t[2];
  int ntasks;
  int batch;
  const int* s;     // per-matrix number of squarings (may be null when no task is predicated)
  const int* sym = nullptr;   // per-matrix symmetry class: 0 general, 1 skew-Hermitian, 2 Hermitian (null: all general)
  // stream-K scratch (set by launch() when it picks the stream-K grid)
  cx<T>* partial = nullptr;       // [2 * CTA + head/tail][32 * NTHREADS] accumulator images
  unsigned* counters = nullptr;   // [image slot] ready flags, zero between launches (the reader resets them);
                                  // [kSplitSlots] = CTA ticket, [kSplitSlots + 1] = CTAs done (the last one resets both)
};
constexpr int kSplitSlots = 148 * 8;       // accumulator images / ready flags a stream-K launch may use

__device__ __forceinline__ int swz(int r) { return ((r & 1) << 2) | (r & 2); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// ---- stage loads ---------------------------------------------------------------------------
__device__ __forceinline__ void load_stage(cx<double>* sA, cx<double>* sB, const cx<double>* A, int lda,
                                           const cx<double>* B, int ldb, int m0, int n0, int k0) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int q = 0; q < (BM * BK) / NTHREADS; ++q) {
    const int e = tid + q * NTHREADS, r = e / BK, k = e % BK;
    cp_async16(&sA[r * BK + (k ^ swz(r))], &A[(size_t)(m0 + r) * lda + k0 + k]);
  }
#pragma unroll
  for (int q = 0; q < (BK * BN) / NTHREADS; ++q) {
    const int e = tid + q * NTHREADS, k = e / BN, c = e % BN;
    cp_async16(&sB[k * BN + (c ^ swz(k))], &B[(size_t)(k0 + k) * ldb + n0 + c]);
  }
}
// complex64: 16-byte copies move element PAIRS (the swizzle keeps even-aligned pairs adjacent)
__device__ __forceinline__ void load_stage(cx<float>* sA, cx<float>* sB, const cx<float>* A, int lda,
                                           const cx<float>* B, int ldb, int m0, int n0, int k0) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int q = 0; q < (BM * BK / 2) / NTHREADS; ++q) {
    const int e = tid + q * NTHREADS, r = e / (BK / 2), k = 2 * (e % (BK / 2));
    cp_async16(&sA[r * BK + (k ^ swz(r))], &A[(size_t)(m0 + r) * lda + k0 + k]);
  }
#pragma unroll
  for (int q = 0; q < (BK * BN / 2) / NTHREADS; ++q) {
    const int e = tid + q * NTHREADS, k = e / (BN / 2), c = 2 * (e % (BN / 2));
    cp_async16(&sB[k * BN + (c ^ swz(k))], &B[(size_t)(k0 + k) * ldb + n0 + c]);
  }
}

// ---- per-thread accumulators -----------------------------------------------------------------
template <typename T> struct Acc;
template <> struct Acc<double> {
  double cr[4][4][2], ci[4][4][2];     // 32x32 warp tile = 4x4 MMA tiles, (re, im) planes
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { cr[i][j][0] = cr[i][j][1] = ci[i][j][0] = ci[i][j][1] = 0.0; }
  }
  __device__ __forceinline__ void compute(const cx<double>* sA, const cx<double>* sB) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gid = lane >> 2, tig = lane & 3, wm = warp >> 1, wn = warp & 1;
#pragma unroll
    for (int ks = 0; ks < BK / 4; ++ks) {
      const int k = ks * 4 + tig;
      cx<double> a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { const int r = wm * 32 + i * 8 + gid; a[i] = sA[r * BK + (k ^ swz(r))]; }
#pragma unroll
      for (int j = 0; j < 4; ++j) { const int c = wn * 32 + j * 8 + gid; b[j] = sB[k * BN + (c ^ swz(k))]; }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double nai = -a[i].im;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          dmma(cr[i][j], a[i].re, b[j].re);
          dmma(ci[i][j], a[i].re, b[j].im);
          dmma(cr[i][j], nai, b[j].im);
          dmma(ci[i][j], a[i].im, b[j].re);
        }
      }
    }
  }
  template <typename F> __device__ __forceinline__ void for_each(F f) const {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gid = lane >> 2, tig = lane & 3, wm = warp >> 1, wn = warp & 1;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e)
          f(wm * 32 + i * 8 + gid, wn * 32 + j * 8 + tig * 2 + e, cx<double>(cr[i][j][e], ci[i][j][e]));
  }
  // the same walk with the running element index q = 0 ... 31 (a compile-time constant after unrolling)
  template <typename F> __device__ __forceinline__ void for_each_q(F f) const {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gid = lane >> 2, tig = lane & 3, wm = warp >> 1, wn = warp & 1;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e)
          f((i * 4 + j) * 2 + e, wm * 32 + i * 8 + gid, wn * 32 + j * 8 + tig * 2 + e, cx<double>(cr[i][j][e], ci[i][j][e]));
  }
  // accumulator image: value q of thread t at [q * NTHREADS + t] (coalesced 16-byte accesses)
  __device__ __forceinline__ void park(cx<double>* img) const {
#pragma unroll
    for (int q = 0; q < 32; ++q)
      __stcg(reinterpret_cast<double2*>(img + q * NTHREADS + threadIdx.x),
             make_double2(cr[q >> 3][(q >> 1) & 3][q & 1], ci[q >> 3][(q >> 1) & 3][q & 1]));
  }
  __device__ __forceinline__ void add_parked(const cx<double>* img) {
#pragma unroll
    for (int q = 0; q < 32; ++q) {
      const double2 v = __ldcg(reinterpret_cast<const double2*>(img + q * NTHREADS + threadIdx.x));
      cr[q >> 3][(q >> 1) & 3][q & 1] += v.x; ci[q >> 3][(q >> 1) & 3][q & 1] += v.y;
    }
  }
};
template <> struct Acc<float> {
  // rows ty*8 + i (i < 8); columns tx*2 + (j & 1) + 32 * (j >> 1) (j < 4): two element PAIRS 32 columns
  // apart, so that a half-warp's 16-byte B loads are contiguous; the swizzles are even, so the pairs
  // (k, k+1) of A and (c, c+1) of B stay adjacent in shared memory and load as one LDS.128 each:
  // 6 shared-memory loads per 128 FMAs.
  cx<float> c[8][4];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) c[i][j] = cx<float>(0.f, 0.f);
  }
  __device__ __forceinline__ void compute(const cx<float>* sA, const cx<float>* sB) {
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
    for (int k2 = 0; k2 < BK; k2 += 2) {
      float4 a4[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = ty * 8 + i;
        a4[i] = *reinterpret_cast<const float4*>(&sA[r * BK + (k2 ^ swz(r))]);
      }
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int k = k2 + kk;
        const float4 b01 = *reinterpret_cast<const float4*>(&sB[k * BN + ((tx * 2) ^ swz(k))]);
        const float4 b23 = *reinterpret_cast<const float4*>(&sB[k * BN + ((32 + tx * 2) ^ swz(k))]);
        const cx<float> b0(b01.x, b01.y), b1(b01.z, b01.w), b2(b23.x, b23.y), b3(b23.z, b23.w);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const cx<float> a = kk ? cx<float>(a4[i].z, a4[i].w) : cx<float>(a4[i].x, a4[i].y);
          cfma(c[i][0], a, b0); cfma(c[i][1], a, b1); cfma(c[i][2], a, b2); cfma(c[i][3], a, b3);
        }
      }
    }
  }
  template <typename F> __device__ __forceinline__ void for_each(F f) const {
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) f(ty * 8 + i, tx * 2 + (j & 1) + 32 * (j >> 1), c[i][j]);
  }
  template <typename F> __device__ __forceinline__ void for_each_q(F f) const {
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) f(i * 4 + j, ty * 8 + i, tx * 2 + (j & 1) + 32 * (j >> 1), c[i][j]);
  }
  __device__ __forceinline__ void park(cx<float>* img) const {
#pragma unroll
    for (int q = 0; q < 32; ++q)
      __stcg(reinterpret_cast<float2*>(img + q * NTHREADS + threadIdx.x), make_float2(c[q >> 2][q & 3].re, c[q >> 2][q & 3].im));
  }
  __device__ __forceinline__ void add_parked(const cx<float>* img) {
#pragma unroll
    for (int q = 0; q < 32; ++q) {
      const float2 v = __ldcg(reinterpret_cast<const float2*>(img + q * NTHREADS + threadIdx.x));
      c[q >> 2][q & 3].re += v.x; c[q >> 2][q & 3].im += v.y;
    }
  }
};

// One CTA = one 64x64 output tile (SK = false, grid (tiles_n, tiles_m, ntasks*batch)), or -- stream-K,
// SK = true, 1-D grid of persistent CTAs -- an equal share of ALL k iterations of the launch, cut where
// it falls: a CTA then owns the tail of one tile, some whole tiles and the head of another.  Heads are
// parked as accumulator images; the CTA that owns a tile's tail adds them onto its own accumulators and
// runs the epilogue.  Stream-K is what keeps 148 SMs busy when a launch has only one or two waves of
// tiles (one 1024^3 product = 256 tiles).
template <typename T, bool SK>
__global__ void __launch_bounds__(NTHREADS, 2) gemm_kernel(const __grid_constant__ Launch<T> L) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<T>* smem = reinterpret_cast<cx<T>*>(smem_raw);
  auto stage_a = [&](int st) { return smem + (size_t)st * (BM * BK + BK * BN); };
  auto stage_b = [&](int st) { return smem + (size_t)st * (BM * BK + BK * BN) + BM * BK; };

  // k iterations [ib, ie) of tile (tm, tn) of matrix b of task t, into acc
  auto mainloop = [&](Acc<T>& acc, const Task<T>& t, int b, int m0, int n0, int ib, int ie) {
    const int KT = t.K / BK;
    const int total = ie - ib;
    auto issue = [&](int itl) {
      const int it = itl + ib;
      const int term = it / KT, kt = it - term * KT;
      const cx<T>* A = (term ? t.A1 + (size_t)b * t.sA1 : t.A0 + (size_t)b * t.sA0);
      const cx<T>* B = (term ? t.B1 + (size_t)b * t.sB1 : t.B0 + (size_t)b * t.sB0);
      load_stage(stage_a(itl % STAGES), stage_b(itl % STAGES), A, term ? t.lda1 : t.lda0, B,
                 term ? t.ldb1 : t.ldb0, m0, n0, kt * BK);
    };
    acc.zero();
#pragma unroll
    for (int p = 0; p < STAGES - 1; ++p) {
      if (p < total) issue(p);
      cp_async_commit();
    }
    for (int it = 0; it < total; ++it) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      if (it + STAGES - 1 < total) issue(it + STAGES - 1);
      cp_async_commit();
      acc.compute(stage_a(it % STAGES), stage_b(it % STAGES));
    }
    cp_async_wait<0>();
  };

  // mir: 0 = plain; 1 = also write the mirror tile as conj(C1) / conj(C2); 2 = mirror of C1 is conj(C2) and vice versa;
  // 3 = the mirror tile is -conj(C1) / -conj(C2)
  auto epilogue = [&](const Acc<T>& acc, const Task<T>& t, int b, int m0, int n0, int mir) {
    cx<T>* C1 = t.C1 + (size_t)b * t.sC1;
    cx<T>* C2 = t.C2 ? t.C2 + (size_t)b * t.sC2 : nullptr;
    const cx<T>* X0 = t.X0 ? t.X0 + (size_t)b * t.sX0 : nullptr;
    const cx<T>* X1 = t.X1 ? t.X1 + (size_t)b * t.sX1 : nullptr;
    const cx<T>* X2 = t.X2 ? t.X2 + (size_t)b * t.sX2 : nullptr;
    // The linear terms are read in chunks of kEpiChunk elements BEFORE the chunk's stores: X may alias C (the
    // in-place rank-64 update of the blocked inverse, the dual recurrences), so written element by element every
    // load waits behind the previous store -- 32 dependent L2 round trips per tile, which was half the time of a
    // K = 64 launch (DMMA pipe 43 % active).
    constexpr int kEpiChunk = SK ? 4 : 8;                             // the stream-K variant has no registers to spare
#pragma unroll
    for (int base = 0; base < 32; base += kEpiChunk) {
      cx<T> lin[kEpiChunk];
      acc.for_each_q([&](int q, int rl, int cl, cx<T>) {
        if (q < base || q >= base + kEpiChunk) return;
        const int r = m0 + rl, c = n0 + cl;
        cx<T> l(r == c ? t.xI : T(0), T(0));
        if (X0) { const cx<T> x = X0[(size_t)r * t.ldx0 + c]; l.re = fma(t.x0, x.re, l.re); l.im = fma(t.x0, x.im, l.im); }
        if (X1) { const cx<T> x = X1[(size_t)r * t.ldx1 + c]; l.re = fma(t.x1, x.re, l.re); l.im = fma(t.x1, x.im, l.im); }
        if (X2) { const cx<T> x = X2[(size_t)r * t.ldx2 + c]; l.re = fma(t.x2, x.re, l.re); l.im = fma(t.x2, x.im, l.im); }
        lin[q - base] = l;
      });
      acc.for_each_q([&](int q, int rl, int cl, cx<T> a) {
        if (q < base || q >= base + kEpiChunk) return;
        const int r = m0 + rl, c = n0 + cl;
        const cx<T> l = lin[q - base];
        const cx<T> o1(fma(t.al1, a.re, t.sg1 * l.re), fma(t.al1, a.im, t.sg1 * l.im));
        C1[(size_t)r * t.ldc1 + c] = o1;
        cx<T> o2 = o1;
        if (C2) {
          o2 = cx<T>(fma(t.al2, a.re, t.sg2 * l.re), fma(t.al2, a.im, t.sg2 * l.im));
          C2[(size_t)r * t.ldc2 + c] = o2;
        }
        if (mir == 3) {
          C1[(size_t)c * t.ldc1 + r] = cx<T>(-o1.re, o1.im);
          if (C2) C2[(size_t)c * t.ldc2 + r] = cx<T>(-o2.re, o2.im);
        } else if (mir) {
          C1[(size_t)c * t.ldc1 + r] = conj(mir == 2 ? o2 : o1);
          if (C2) C2[(size_t)c * t.ldc2 + r] = conj(mir == 2 ? o1 : o2);
        }
      });
    }
  };
  // structure mode of task t for matrix b: 0 none, 1 self-mirror, 2 pair-mirror
  auto herm_mode = [&](const Task<T>& t, int b) {
    if (!t.herm || !L.sym) return 0;
    const int sy = L.sym[b];
    if (!sy) return 0;
    if (t.herm == 3) return sy == 1 ? 3 : 0;
    return (t.herm == 2 && sy == 1) ? 2 : 1;
  };

  Acc<T> acc;
  if (!SK) {
    const int task_id = blockIdx.z / L.batch;
    const int b = blockIdx.z - task_id * L.batch;
    const Task<T>& t = L.t[task_id];
    const int tm = blockIdx.y, tn = blockIdx.x;
    if (tm * BM >= t.M || tn * BN >= t.N) return;
    if (tm == t.skip_tm || tn == t.skip_tn) return;
    if (t.pred_it >= 0 && !(t.pred_it < L.s[b])) return;
    const int hm = herm_mode(t, b);
    if (hm && tn < tm) return;                                               // the mirror of tile (tn, tm) covers it
    mainloop(acc, t, b, tm * BM, tn * BN, 0, (t.K / BK) * (t.A1 ? 2 : 1));
    epilogue(acc, t, b, tm * BM, tn * BN, tn > tm ? hm : 0);
    return;
  }

  // ---- stream-K: all tasks share one shape (launch() checks), so a tile index decodes uniformly.
  // Share c owns global k iterations [I c / G, I (c+1) / G) and walks them BACKWARDS: first the head of the
  // last tile it touches (parked at once, early in its life), then its whole tiles, last the tail of the
  // first tile -- which it finishes: by then the shares below it (c-1, ...) have long parked that tile's
  // earlier parts, so the flag wait is a formality, and since a share only ever waits on lower-numbered
  // shares there is no cycle.  Images are added in share order onto the finisher's own accumulators:
  // bit-reproducible.
  // Forward progress does not lean on the hardware's CTA dispatch order or on all G CTAs being co-resident (another
  // stream, an NCCL kernel or MIG may hold SMs): a CTA's share index is a TICKET drawn when it starts running, so
  // every share a CTA can wait on belongs to a CTA that is already executing.  What a share computes depends on its
  // index only, so the result is the same whichever CTA draws it.
  const Task<T>& t0 = L.t[0];
  const int tnc = t0.N / BN, tmc = t0.M / BM;
  const long long ipt = (long long)(t0.K / BK) * (t0.A1 ? 2 : 1);            // k iterations per tile
  // tiles per (task, matrix): all of them, or the upper triangle when the structure hint applies
  const int nz = L.batch * L.ntasks, full = tnc * tmc, tri = tmc * (tmc + 1) / 2;
  auto tiles_of = [&](int zz) { const int tk = zz / L.batch; return herm_mode(L.t[tk], zz - tk * L.batch) ? tri : full; };
  long long ntiles = 0;
  for (int zz = 0; zz < nz; ++zz) ntiles += tiles_of(zz);
  const long long I = ipt * ntiles;
  __shared__ unsigned ticket_s;
  if (threadIdx.x == 0) ticket_s = atomicAdd(L.counters + kSplitSlots, 1u);
  __syncthreads();
  const long long G = gridDim.x, c = ticket_s;
  const long long g0 = I * c / G;
  constexpr size_t IMG = (size_t)32 * NTHREADS;
  for (long long g_hi = I * (c + 1) / G; g_hi > g0;) {
    const long long tile = (g_hi - 1) / ipt, tbeg = tile * ipt;
    const int ib = (int)((g0 > tbeg ? g0 : tbeg) - tbeg);
    const int ie = (int)(g_hi - tbeg);
    g_hi = tbeg + ib;
    int zz = 0, tl = (int)tile;
    for (int cnt = tiles_of(0); tl >= cnt; cnt = tiles_of(zz)) { tl -= cnt; ++zz; }
    const int task_id = zz / L.batch, b = zz - task_id * L.batch;
    const Task<T>& t = L.t[task_id];
    const int hm = herm_mode(t, b);
    int tm, tn;
    if (hm) { tm = 0; for (int w = tmc; tl >= w; --w) { tl -= w; ++tm; } tn = tm + tl; }   // row tm of the upper triangle has tmc - tm tiles
    else { tm = tl / tnc; tn = tl - tm * tnc; }
    const int mir = tn > tm ? hm : 0;
    if (t.pred_it >= 0 && !(t.pred_it < L.s[b])) continue;                   // every CTA on this tile skips alike
    mainloop(acc, t, b, tm * BM, tn * BN, ib, ie);
    if (ie != (int)ipt) {
      // not the end of the tile: park (slot 0: this CTA holds the tile's head, slot 1: a middle part)
      const size_t slot = (size_t)(2 * c + (ib == 0 ? 0 : 1));
      acc.park(L.partial + slot * IMG);
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) { volatile unsigned* f = L.counters + slot; *f = 1u; }
    } else {
      if (ib != 0) {
        // CTA q holds global iteration x iff q = ceil((x+1) G / I) - 1
        const long long c_first = ((tbeg + 1) * G + I - 1) / I - 1;
        for (long long q = c_first; q < c; ++q) {
          const size_t slot = (size_t)(2 * q + (I * q / G <= tbeg ? 0 : 1));
          if (threadIdx.x == 0) {
            volatile unsigned* f = L.counters + slot;
            while (*f == 0u) {}
            *f = 0u;                                                         // ready for the next launch
          }
          __syncthreads();
          __threadfence();
          acc.add_parked(L.partial + slot * IMG);
        }
      }
      epilogue(acc, t, b, tm * BM, tn * BN, mir);
    }
    __syncthreads();                                                         // smem stages are reused by the next segment
  }
  // the last CTA out re-arms the ticket for the next launch on this stream (launches on one stream do not overlap)
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(L.counters + kSplitSlots + 1, 1u) == (unsigned)G - 1u) {
      L.counters[kSplitSlots] = 0u; L.counters[kSplitSlots + 1] = 0u;
      __threadfence();
    }
  }
}

// ---- stream-K scratch: one per (device, stream), allocated on first use, released by qg_release_scratch --------
struct SplitScratch { void* partial; unsigned* counters; };
int split_scratch(cudaStream_t st, size_t image_bytes, SplitScratch* out);   // api.cu
int device_sm_count();                                                       // api.cu (cached per device)

// Stream-K grid for a launch of `tiles` tiles with `ipt` k iterations each, or 0 for the plain tile grid.
// Two CTAs share an SM's tensor pipe, so a plain launch takes ceil(tiles / (2 SMs)) "waves"; stream-K pays when
// the last wave is poorly filled and the shares stay long enough to amortise the parking.
inline int choose_streamk(long long tiles, long long ipt) {
  long long slots = 2 * (long long)device_sm_count();
  if (2 * slots > kSplitSlots) slots = kSplitSlots / 2;                      // two image slots per share
  if (tiles >= 4 * slots || tiles >= kSplitSlots || ipt < 16) return 0;
  const double waves = (double)tiles / slots;
  const double eff = waves / (double)((tiles + slots - 1) / slots);
  if (eff > 0.92) return 0;
  long long G = slots;
  while (G > 1 && tiles * ipt / G < 8) G /= 2;                               // at least 8 iterations per CTA
  return G >= 2 ? (int)G : 0;
}

// complex64 on the tcgen05 tensor cores (gemm_tc.cu): launch() returns kNotTaken when the launch does not qualify
namespace tc {
constexpr int kNotTaken = -77001;
int launch(const Launch<float>& L, cudaStream_t st);
bool enabled();
bool usable();
int set_enabled(int on);
}  // namespace tc

template <typename T>
inline int launch(const Launch<T>& Lin, cudaStream_t st) {
  if constexpr (std::is_same<T, float>::value) {
    const int rc = tc::launch(Lin, st);
    if (rc != tc::kNotTaken) return rc;
  }
  Launch<T> L = Lin;
  int maxM = 0, maxN = 0;
  for (int i = 0; i < L.ntasks; ++i) { if (L.t[i].M > maxM) maxM = L.t[i].M; if (L.t[i].N > maxN) maxN = L.t[i].N; }
  if (maxM <= 0 || maxN <= 0 || L.batch <= 0) return QG_OK;
  const size_t smem = (size_t)STAGES * (BM * BK + BK * BN) * sizeof(cx<T>);
  static std::atomic<unsigned long long> attr_done{0};
  if (first_use_on_device(attr_done)) {
    QG_RETURN_IF_CUDA(cudaFuncSetAttribute(gemm_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QG_RETURN_IF_CUDA(cudaFuncSetAttribute(gemm_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  // stream-K only for single-shape launches without skipped tile rows/columns (a predicated matrix drops
  // out of every CTA that touches it alike, before the counter)
  bool uniform = true;
  for (int i = 0; i < L.ntasks; ++i)
    uniform = uniform && L.t[i].M == maxM && L.t[i].N == maxN && L.t[i].skip_tm < 0 && L.t[i].skip_tn < 0 &&
              L.t[i].K == L.t[0].K && (L.t[i].A1 != nullptr) == (L.t[0].A1 != nullptr);
  if (uniform) {
    const long long tiles = (long long)(maxN / BN) * (maxM / BM) * L.ntasks * L.batch;
    const long long ipt = (long long)(L.t[0].K / BK) * (L.t[0].A1 ? 2 : 1);
    int G = choose_streamk(tiles, ipt);
    static const int forced = [] { const char* e = getenv("QG_STREAMK"); return e ? atoi(e) : -1; }();   // tuning aid
    if (forced == 0) G = 0;
    if (forced > 0 && tiles < kSplitSlots && 2 * forced <= kSplitSlots && tiles * ipt >= forced) G = forced;
    if (G > 0) {
      SplitScratch sc;
      int rc = split_scratch(st, (size_t)32 * NTHREADS * sizeof(cx<double>), &sc);
      if (rc) return rc;
      L.partial = (cx<T>*)sc.partial; L.counters = sc.counters;
      gemm_kernel<T, true><<<dim3((unsigned)G), NTHREADS, smem, st>>>(L);
      QG_CHECK_LAUNCH();
      return QG_OK;
    }
  }
  const long long gz = (long long)L.ntasks * L.batch;
  if (gz > 65535) return QG_ERR_ARG;
  dim3 grid(maxN / BN, maxM / BM, (unsigned)gz);
  gemm_kernel<T, false><<<grid, NTHREADS, smem, st>>>(L);
  QG_CHECK_LAUNCH();
  return QG_OK;
}

}  // namespace gemm
}  // namespace qg

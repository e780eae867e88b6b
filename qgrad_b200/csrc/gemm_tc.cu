// complex64 GEMM on the 5th-generation tensor cores (tcgen05.mma.kind::tf32, accumulators in TMEM, operands staged by
// TMA), carrying the same Task as gemm.cuh:  acc = A0.B0 (+ A1.B1);  C1 = al1 acc + sg1 lin;  C2 = al2 acc + sg2 lin.
//
// float32 accuracy from TF32 products (split a = hi + lo, both TF32, round-to-nearest):
//     a.b  ~=  hi_a hi_b + lo_a hi_b + hi_a lo_b                      (the dropped lo_a lo_b is 2^-22 of the product)
// and a complex product is four real ones, so one k-step of 8 is 12 MMAs of 128 x 128 x 8:
//     Cr += Ar.Br - Ai.Bi        Ci += Ar.Bi + Ai.Br                  (the minus sign is the descriptor's negate-A bit)
// The tensor core adds into its FP32 accumulator with TRUNCATION: every MMA shrinks the running sum by ~0.36 ulp on
// average, a COHERENT relative error (C -> (1 - delta) C) that a chain of squarings doubles at every step while its
// rounding noise only grows by sqrt 2 -- measured on B200 (tools/tc_gemm_check.py): with everything accumulated in TMEM
// over 32 k, delta = 3e-7 per product and 16 chained squarings of a unitary end 19x further off than the FMA kernel's.
// So (Ootomo & Yokota's correction, in TMEM terms):
//   * the LEADING products (hi hi) go to their own TMEM accumulator (Cr | Ci, 256 columns) that only ever holds a
//     short partial sum -- one 16-k stage = four dependent MMAs -- and is then drained by eight epilogue warps into
//     FP32 registers with round-to-nearest adds, while the tensor core works through the stage's correction products;
//   * the CORRECTION products (lo hi, hi lo: 2^-11 of the result, their truncation is noise) accumulate over the whole
//     k range in a second accumulator (256 columns), drained once per tile.
//
// Shared-memory bandwidth is what bounds a kernel like this (every MMA reads both operands: 128 B/clk for 128 x 128 x 8
// steps, plus the conversion traffic; the first version, with both operands in shared memory, ran at 53 % of the tensor pipe
// with the pipe "active" 82 % of the time).  So the A operand never touches shared memory after the raw tile: the converter
// warps write its four planes straight into TENSOR MEMORY (tcgen05.st) and the MMAs read A from there (the .ts form).  TMEM
// then holds, per CTA, the leading accumulators (Cr | Ci), the correction accumulators (Cr | Ci) and two stages of A planes
// (64 columns each): 4 N + 128 = 512 columns at N = 96 -- hence the 128 x 96 complex tile.
//
// One CTA per 128 x 96 complex output tile, 16 warps, four roles:
//     warp 0        TMA producer: raw complex64 tiles of A (128 x 16, 128-byte swizzle) and B (16 x 96) -> shared memory
//     warps 4-7     B converters: raw tile -> planes (-Bi | Br | Bi), hi and lo, in shared memory in the tensor core's K-major
//                   no-swizzle layout ([column/8][k/4][8 columns][4 k] = 128-byte core matrices; B is transposed on the way)
//     warp 1        MMA issuer (one lane): 24 tcgen05.mma per 16-k stage, tcgen05.commit onto the stage / TMEM barriers
//     warps 8-15    accumulators + epilogue: tcgen05.ld the partial sums, add, at the end the linear-combination
//                   epilogue of gemm.cuh (row per thread, 48 complex columns); in the time they would otherwise wait they
//                   convert A: thread = (row, half of the stage's k): raw row -> hi / lo, re / im planes -> TMEM
// Registers are re-balanced between the roles with setmaxnreg: 512 threads start at 128 registers; the TMA / MMA warpgroup
// goes down to 56, the B converters to 72, and the 16384 registers they release take the two accumulator warpgroups up to 192
// (setmaxnreg.inc only ever draws on what the CTA itself has released: a 640-thread variant whose increase counted on the
// SM's unallocated registers hung).
#include <cuda.h>
#include <mutex>
#include "gemm.cuh"

namespace qg {
namespace gemm {
namespace tc {

constexpr int TM = 128, TN = 96;              // complex output tile
constexpr int KS = 16;                        // complex k per stage
constexpr int RAW_STAGES = 4, OP_STAGES = 2;
constexpr int RAW_A_BYTES = TM * KS * 8, RAW_B_BYTES = KS * TN * 8, RAW_BYTES = RAW_A_BYTES + RAW_B_BYTES;   // 32 KB
constexpr int PLANE_BYTES = TN * KS * 4;      // one 96 x 16 float plane of B: 6 KB
constexpr int OP_BYTES = 6 * PLANE_BYTES;     // B: (-ih | rh | ih) (-il | rl | il) = 36 KB (the planes of A live in TMEM)
constexpr int SMEM_BYTES = 1024 + OP_STAGES * OP_BYTES + RAW_STAGES * RAW_BYTES + 256;
constexpr int NTHREADS = 512;
constexpr int kConvThreads = 128, kEpiThreads = 256;        // B converters; accumulator warps (which also convert A)
constexpr int kStageThreads = kConvThreads + kEpiThreads;   // arrivals that complete a stage of operands / free a raw stage
constexpr int TMEM_COLS = 512;                // leading (Cr | Ci) 192 + corrections (Cr | Ci) 192 + 2 stages of A planes 128
constexpr int COL_RE = 0, COL_IM = TN, COL_CRE = 2 * TN, COL_CIM = 3 * TN, COL_A = 4 * TN, A_STAGE_COLS = 4 * KS;
static_assert(COL_A + OP_STAGES * A_STAGE_COLS <= TMEM_COLS, "TMEM budget");

// ---- PTX wrappers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LAB_DONE;\n"
      "bra LAB_WAIT;\n"
      "LAB_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// one lane of a converged warp (the compiler keeps warp-uniform operands in uniform registers only in convergent code: an
// `if (lane == 0)` around the issue loop made every tcgen05.mma a waterfall loop of ~12 instructions)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.b32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A from tensor memory (lane = row, one column per TF32 element), B from shared memory
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
      "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
      "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tc_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
// a = hi + lo with hi = a rounded to TF32 (to nearest, ties away, like cvt.rna.tf32.f32 -- which ptxas expands to ~7
// instructions per value and made the four converter warps the bottleneck): rounding on the bit pattern is an add and a mask.
// lo = a - hi is exact in float32 and goes to the tensor core as it is: the hardware ignores the low 13 bits of an operand,
// i.e. truncates lo to 11 bits -- an error below 2^-22 |a| whose sign follows lo's, not a's (no coherent shrink).
__device__ __forceinline__ void split_tf32(float a, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(a) + 0x1000u) & 0xFFFFE000u;
  lo = __float_as_uint(a - __uint_as_float(hi));
}
// K-major, no swizzle: core matrix = 8 rows x 16 bytes, contiguous (128 B).  LBO = distance between the two core matrices
// of one MMA along k, SBO = distance between 8-row groups (cute::UMMA::SmemDescriptor, version 1 = Blackwell).
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// kind::tf32, D = F32, A and B TF32 K-major, M = 128, N = 192 = (real | imaginary) columns (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t instr_desc(bool neg_a) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((neg_a ? 1u : 0u) << 13) | ((uint32_t)((2 * TN) >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}

struct Params {
  Task<float> t;
  int batch;
  const int* s;
  const int* sym;     // per-matrix symmetry class (gemm.cuh), may be null
  int kstages;        // 16-k stages per term
  int nterms;         // 1 or 2
  float debias;       // mean relative truncation loss of a partial sum (see the epilogue)
};

template <int CH>   // stages per TMEM partial sum of the leading products (1; 2 as an experiment, QG_TC_CHUNK)
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapB0,
               const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapB1,
               const __grid_constant__ Params P) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const Task<float>& t = P.t;
  const int b = blockIdx.z;
  if (t.pred_it >= 0 && !(t.pred_it < P.s[b])) return;                       // whole CTA alike
  // structure hint (gemm.cuh): 0 none; 1 mirror = conj; 2 mirror of C1 = conj(C2) and vice versa; 3 mirror = -conj
  int hm = 0;
  if (t.herm && P.sym) {
    const int sy = P.sym[b];
    if (sy) hm = t.herm == 3 ? (sy == 1 ? 3 : 0) : ((t.herm == 2 && sy == 1) ? 2 : 1);
  }
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  if (m0 == t.skip_r0) return;                                               // (the 128-wide bands of the blocked inverse)
  if (t.skip_c0 >= 0 && n0 >= t.skip_c0 && n0 + TN <= t.skip_c0 + 128) return;
  // with a structure hint only elements on or above the diagonal are written directly; each strictly-upper element also
  // writes its mirror image.  A tile wholly below the diagonal has nothing to do.
  if (hm && n0 + TN - 1 < m0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t op_base = base;                                             // OP_STAGES x 24 KB (planes of B)
  const uint32_t raw_base = base + 1024 * ((OP_STAGES * OP_BYTES + 1023) / 1024);        // RAW_STAGES x (A 16 KB | B 12 KB), 1024-aligned
  const uint32_t bar_base = raw_base + RAW_STAGES * RAW_BYTES;
  auto full_raw = [&](int i) { return bar_base + 8u * i; };
  auto empty_raw = [&](int i) { return bar_base + 8u * (RAW_STAGES + i); };
  auto full_op = [&](int i) { return bar_base + 8u * (2 * RAW_STAGES + i); };
  auto empty_op = [&](int i) { return bar_base + 8u * (2 * RAW_STAGES + OP_STAGES + i); };
  const uint32_t re_full = bar_base + 8u * (2 * RAW_STAGES + 2 * OP_STAGES);
  const uint32_t re_empty = bar_base + 8u * (2 * RAW_STAGES + 2 * OP_STAGES + 1);
  const uint32_t corr_full = bar_base + 8u * (2 * RAW_STAGES + 2 * OP_STAGES + 2);
  const uint32_t tmem_slot = bar_base + 8u * (2 * RAW_STAGES + 2 * OP_STAGES + 4);   // (4 bytes)
  unsigned char* gen_base = smem_raw + (base - smem_u32(smem_raw));          // generic pointer to `base`

  if (threadIdx.x == 0) {
    for (int i = 0; i < RAW_STAGES; ++i) { mbar_init(full_raw(i), 1); mbar_init(empty_raw(i), kStageThreads); }
    for (int i = 0; i < OP_STAGES; ++i) { mbar_init(full_op(i), kStageThreads); mbar_init(empty_op(i), 1); }
    mbar_init(re_full, 1); mbar_init(re_empty, kEpiThreads);
    mbar_init(corr_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(tmem_slot), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - base));

  const int NS = P.kstages * P.nterms;                                       // stages in all
  const int NCH = (NS + CH - 1) / CH;                                        // TMEM partial sums of the leading products

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;\n");
    if (warp == 0) {
      // ---- TMA producer (whole warp in the loop, one elected lane issues)
      int term = 0, k0 = 0;
      for (int s = 0; s < NS; ++s, k0 += KS) {
        if (k0 == P.kstages * KS) { k0 = 0; term = 1; }
        const int slot = s % RAW_STAGES, ph = (s / RAW_STAGES) & 1;
        mbar_wait(empty_raw(slot), ph ^ 1);
        const uint32_t dst = raw_base + slot * RAW_BYTES;
        if (elect_one()) {
          mbar_expect_tx(full_raw(slot), RAW_BYTES);
          tma_load_3d(dst, term ? &mapA1 : &mapA0, full_raw(slot), 2 * k0, m0, b);
          tma_load_3d(dst + RAW_A_BYTES, term ? &mapB1 : &mapB0, full_raw(slot), 2 * n0, k0, b);
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      // ---- MMA issuer
      // One MMA covers the real AND the imaginary columns: with the planes of B laid out as (-Bi | Br | Bi), rows 96 ... 287 are
      // the operand (Br | Bi) and rows 0 ... 191 the operand (-Bi | Br):
      //     (Cr | Ci) += Ar (Br | Bi) + Ai (-Bi | Br)
      // N = 192 per instruction (a 96-column MMA was measured at the cost of a 128-column one).
      constexpr uint32_t ID = instr_desc(false);
      constexpr uint32_t LBO = 128, SBO = 512;
      const uint32_t d_main = tmem + COL_RE, d_corr = tmem + COL_CRE;
#pragma unroll 1
      for (int s = 0; s < NS; ++s) {
        const int slot = s % OP_STAGES, ph = (s / OP_STAGES) & 1;
        const int c = s / CH, sc = s - c * CH;
        mbar_wait(full_op(slot), ph);
        const bool last = sc == CH - 1 || s == NS - 1;
        if (sc == 0) mbar_wait(re_empty, (c & 1) ^ 1);                       // the previous partial sum is in registers
        tc_fence_after();
        // A plane p (0 rh, 1 rl, 2 ih, 3 il) in TMEM columns; B window w (0: (-Bi | Br), 1: (Br | Bi)) of the hi (h = 0) or lo planes
        const uint32_t ta = tmem + COL_A + slot * A_STAGE_COLS;
        const uint64_t db = smem_desc(op_base + slot * OP_BYTES, LBO, SBO);
        auto A_ = [&](int p, int ks) { return ta + (uint32_t)(p * KS + ks * 8); };
        auto B_ = [&](int w, int h, int ks) { return db + (uint64_t)(((3 * h + w) * PLANE_BYTES + ks * 2 * (int)LBO) >> 4); };
        if (elect_one()) {
          // leading products first: their accumulator is handed to the epilogue warps, which drain it under the corrections
#pragma unroll
          for (int ks = 0; ks < KS / 8; ++ks) {
            tc_mma_tf32_ts(d_main, A_(0, ks), B_(1, 0, ks), ID, (sc == 0 && ks == 0) ? 0u : 1u);
            tc_mma_tf32_ts(d_main, A_(2, ks), B_(0, 0, ks), ID, 1u);
          }
          if (last) tc_commit(re_full);
#pragma unroll
          for (int ks = 0; ks < KS / 8; ++ks) {
            tc_mma_tf32_ts(d_corr, A_(1, ks), B_(1, 0, ks), ID, (s == 0 && ks == 0) ? 0u : 1u);
            tc_mma_tf32_ts(d_corr, A_(3, ks), B_(0, 0, ks), ID, 1u);
            tc_mma_tf32_ts(d_corr, A_(0, ks), B_(1, 1, ks), ID, 1u);
            tc_mma_tf32_ts(d_corr, A_(2, ks), B_(0, 1, ks), ID, 1u);
          }
          tc_commit(empty_op(slot));                                         // the stage's planes are free once these MMAs retire
          if (s == NS - 1) tc_commit(corr_full);
        }
        __syncwarp();
      }
    }
  } else if (warp < 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;\n");
    // ---- B converters
    const int wc = warp & 3;
    for (int s = 0; s < NS; ++s) {
      const int rslot = s % RAW_STAGES, rph = (s / RAW_STAGES) & 1;
      const int oslot = s % OP_STAGES, oph = (s / OP_STAGES) & 1;
      mbar_wait(full_raw(rslot), rph);
      mbar_wait(empty_op(oslot), oph ^ 1);
      const unsigned char* rawB = gen_base + (raw_base - base) + rslot * RAW_BYTES + RAW_A_BYTES;
      unsigned char* opB = gen_base + oslot * OP_BYTES;
      // B: item (column n, k-core j): transposed on the way (k contiguous per column)
#pragma unroll
      for (int it = 0; it < TN / 32; ++it) {
        const int n = it * 32 + lane, j = wc;
        uint32_t rh[4], rl[4], ih[4], il[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 v = *reinterpret_cast<const float2*>(rawB + (4 * j + q) * (TN * 8) + n * 8);
          split_tf32(v.x, rh[q], rl[q]);
          split_tf32(v.y, ih[q], il[q]);
        }
        const int off = (n >> 3) * 512 + j * 128 + (n & 7) * 16;
        constexpr uint32_t SG = 0x80000000u;
        *reinterpret_cast<uint4*>(opB + off) = make_uint4(ih[0] ^ SG, ih[1] ^ SG, ih[2] ^ SG, ih[3] ^ SG);
        *reinterpret_cast<uint4*>(opB + PLANE_BYTES + off) = make_uint4(rh[0], rh[1], rh[2], rh[3]);
        *reinterpret_cast<uint4*>(opB + 2 * PLANE_BYTES + off) = make_uint4(ih[0], ih[1], ih[2], ih[3]);
        *reinterpret_cast<uint4*>(opB + 3 * PLANE_BYTES + off) = make_uint4(il[0] ^ SG, il[1] ^ SG, il[2] ^ SG, il[3] ^ SG);
        *reinterpret_cast<uint4*>(opB + 4 * PLANE_BYTES + off) = make_uint4(rl[0], rl[1], rl[2], rl[3]);
        *reinterpret_cast<uint4*>(opB + 5 * PLANE_BYTES + off) = make_uint4(il[0], il[1], il[2], il[3]);
      }
      fence_proxy_async();                                                   // generic-proxy stores -> the tensor core's async-proxy reads
      mbar_arrive(full_op(oslot));
      mbar_arrive(empty_raw(rslot));
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 192;\n");
    // ---- accumulators + epilogue: thread = (row, half of the tile's columns)
    const int et = threadIdx.x - 256;
    const int quad = warp & 3, half = et >> 7;
    const int rl_ = quad * 32 + lane;                                        // row inside the tile = TMEM lane
    const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
    constexpr int NC = TN / 2;                                               // complex columns per thread
    float accr[NC], acci[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) { accr[i] = 0.f; acci[i] = 0.f; }
    // both tcgen05.ld in flight before the wait (the drain of the leading products sits on the tensor core's critical path)
    auto drain = [&](uint32_t col0, float (&acc)[NC]) {
      uint32_t v0[32], v1[16];
      tc_ld32(lane_addr + col0, v0);
      tc_ld16(lane_addr + col0 + 32, v1);
      tc_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] += __uint_as_float(v0[i]);
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[32 + i] += __uint_as_float(v1[i]);
    };
    auto drain_chunk = [&](int c) {
      mbar_wait(re_full, c & 1);
      tc_fence_after();
      drain(COL_RE + half * NC, accr);
      drain(COL_IM + half * NC, acci);
      tc_fence_before();
      mbar_arrive(re_empty);
    };
    // These eight warps wait most of a stage for the next partial sum: they also convert A.  Thread = (row = TMEM lane, half of
    // the stage's k): raw row (128-byte swizzle: the eight rows of a quarter-warp hit eight different 16-byte columns) -> hi / lo,
    // re / im -> four 8-column stores into the stage's A planes.  Per stage: planes of stage s first, then the partial sum of
    // stage s - 1 (long complete by then), so the tensor core never waits for either.
    for (int s = 0; s < NS; ++s) {
      const int rslot = s % RAW_STAGES, rph = (s / RAW_STAGES) & 1;
      const int oslot = s % OP_STAGES, oph = (s / OP_STAGES) & 1;
      mbar_wait(full_raw(rslot), rph);
      mbar_wait(empty_op(oslot), oph ^ 1);
      tc_fence_after();
      {
        const unsigned char* rawA = gen_base + (raw_base - base) + rslot * RAW_BYTES;
        float4 v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = *reinterpret_cast<const float4*>(rawA + rl_ * 128 + (((4 * half + q) ^ (rl_ & 7)) << 4));
        const uint32_t ta = lane_addr + COL_A + oslot * A_STAGE_COLS + 8 * half;
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) { split_tf32(v[q].x, hi[2 * q], lo[2 * q]); split_tf32(v[q].z, hi[2 * q + 1], lo[2 * q + 1]); }
        tc_st8(ta, hi); tc_st8(ta + KS, lo);                                 // planes rh, rl
#pragma unroll
        for (int q = 0; q < 4; ++q) { split_tf32(v[q].y, hi[2 * q], lo[2 * q]); split_tf32(v[q].w, hi[2 * q + 1], lo[2 * q + 1]); }
        tc_st8(ta + 2 * KS, hi); tc_st8(ta + 3 * KS, lo);                    // planes ih, il
      }
      tc_st_wait();
      tc_fence_before();
      mbar_arrive(full_op(oslot));
      mbar_arrive(empty_raw(rslot));
      if (s > 0 && (s - 1) % CH == CH - 1) drain_chunk((s - 1) / CH);
    }
    drain_chunk(NCH - 1);
    // The tensor core adds into its accumulator with truncation: each of the 4 n dependent MMAs behind a partial sum of n
    // stages shrinks the running sum by 3.4e-8 of itself on average.  Measured on B200 on dense operands
    // (tools/tc_gemm_check.py, <C - C_ref, C_ref> / <C_ref, C_ref>): -8.64e-8 at n = 1, -1.52e-7 at n = 2, the same to three
    // digits for every shape.  A chain of squarings doubles a coherent error at every step, so its mean is put back here.
    {
      const float kappa = P.debias;
#pragma unroll
      for (int i = 0; i < NC; ++i) { accr[i] = fmaf(accr[i], kappa, accr[i]); acci[i] = fmaf(acci[i], kappa, acci[i]); }
    }
    mbar_wait(corr_full, 0);
    tc_fence_after();
    drain(COL_CRE + half * NC, accr);
    drain(COL_CIM + half * NC, acci);
    // linear-combination epilogue (gemm.cuh), two complex columns (16 bytes) at a time
    const int r = m0 + rl_;
    if (r < t.M) {
      float* C1 = reinterpret_cast<float*>(t.C1 + (size_t)b * t.sC1 + (size_t)r * t.ldc1);
      float* C2 = t.C2 ? reinterpret_cast<float*>(t.C2 + (size_t)b * t.sC2 + (size_t)r * t.ldc2) : nullptr;
      const float* X0 = t.X0 ? reinterpret_cast<const float*>(t.X0 + (size_t)b * t.sX0 + (size_t)r * t.ldx0) : nullptr;
      const float* X1 = t.X1 ? reinterpret_cast<const float*>(t.X1 + (size_t)b * t.sX1 + (size_t)r * t.ldx1) : nullptr;
      const float* X2 = t.X2 ? reinterpret_cast<const float*>(t.X2 + (size_t)b * t.sX2 + (size_t)r * t.ldx2) : nullptr;
#pragma unroll
      for (int i = 0; i < NC; i += 2) {
        const int cc = n0 + half * NC + i;
        if (cc >= t.N) break;
        if (hm && cc + 1 < r) continue;                                      // strictly below the diagonal: a mirror image writes it
        if (t.skip_c0 >= 0 && cc >= t.skip_c0 && cc < t.skip_c0 + 128) continue;
        float4 l = make_float4(cc == r ? t.xI : 0.f, 0.f, cc + 1 == r ? t.xI : 0.f, 0.f);
        if (X0) { const float4 x = *reinterpret_cast<const float4*>(X0 + 2 * cc); l.x = fmaf(t.x0, x.x, l.x); l.y = fmaf(t.x0, x.y, l.y); l.z = fmaf(t.x0, x.z, l.z); l.w = fmaf(t.x0, x.w, l.w); }
        if (X1) { const float4 x = *reinterpret_cast<const float4*>(X1 + 2 * cc); l.x = fmaf(t.x1, x.x, l.x); l.y = fmaf(t.x1, x.y, l.y); l.z = fmaf(t.x1, x.z, l.z); l.w = fmaf(t.x1, x.w, l.w); }
        if (X2) { const float4 x = *reinterpret_cast<const float4*>(X2 + 2 * cc); l.x = fmaf(t.x2, x.x, l.x); l.y = fmaf(t.x2, x.y, l.y); l.z = fmaf(t.x2, x.z, l.z); l.w = fmaf(t.x2, x.w, l.w); }
        const float4 o1 = make_float4(fmaf(t.al1, accr[i], t.sg1 * l.x), fmaf(t.al1, acci[i], t.sg1 * l.y),
                                      fmaf(t.al1, accr[i + 1], t.sg1 * l.z), fmaf(t.al1, acci[i + 1], t.sg1 * l.w));
        float4 o2 = o1;
        if (C2) o2 = make_float4(fmaf(t.al2, accr[i], t.sg2 * l.x), fmaf(t.al2, acci[i], t.sg2 * l.y),
                                 fmaf(t.al2, accr[i + 1], t.sg2 * l.z), fmaf(t.al2, acci[i + 1], t.sg2 * l.w));
        const bool lo_ok = !hm || cc >= r;                                   // (r, cc) is on or above the diagonal; (r, cc + 1) always is here
        if (lo_ok) {
          *reinterpret_cast<float4*>(C1 + 2 * cc) = o1;
          if (C2) *reinterpret_cast<float4*>(C2 + 2 * cc) = o2;
        } else {
          *reinterpret_cast<float2*>(C1 + 2 * cc + 2) = make_float2(o1.z, o1.w);
          if (C2) *reinterpret_cast<float2*>(C2 + 2 * cc + 2) = make_float2(o2.z, o2.w);
        }
        if (hm) {
          // strictly-upper element (r, c) -> (c, r): a warp's 32 rows are 256 contiguous bytes of row c
          const float sgn = hm == 3 ? -1.f : 1.f;
          const float4 w1 = hm == 2 ? o2 : o1, w2 = hm == 2 ? o1 : o2;
          float2* M1 = reinterpret_cast<float2*>(t.C1 + (size_t)b * t.sC1 + (size_t)cc * t.ldc1 + r);
          float2* M2 = C2 ? reinterpret_cast<float2*>(t.C2 + (size_t)b * t.sC2 + (size_t)cc * t.ldc2 + r) : nullptr;
          if (cc > r) { M1[0] = make_float2(sgn * w1.x, -sgn * w1.y); if (M2) M2[0] = make_float2(sgn * w2.x, -sgn * w2.y); }
          if (cc + 1 > r) { M1[t.ldc1] = make_float2(sgn * w1.z, -sgn * w1.w); if (M2) M2[t.ldc2] = make_float2(sgn * w2.z, -sgn * w2.w); }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

// float32 view of a batch of complex64 matrices: dims (2 cols, rows, batch)
static bool make_map(CUtensorMap* m, const void* ptr, long long cols, long long rows, long long batch, long long ld,
                     long long stride, int box_inner, int box_rows, CUtensorMapSwizzle sw) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  if (batch <= 1) { batch = 1; stride = ld * rows; }
  if ((ld & 1) || (stride & 1) || stride <= 0 || ((uintptr_t)ptr & 15)) return false;
  cuuint64_t dims[3] = {(cuuint64_t)(2 * cols), (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 8, (cuuint64_t)stride * 8};
  cuuint32_t box[3] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// 1 = take the tensor-core path when a launch qualifies, 0 = never (QG_C64_TENSOR, qg_set_c64_tensor)
static std::atomic<int> g_enabled{-1};
bool enabled() {
  int v = g_enabled.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("QG_C64_TENSOR");
    v = (e && *e == '0') ? 0 : 1;
    g_enabled.store(v, std::memory_order_relaxed);
  }
  return v != 0;
}
bool usable() { return enabled() && encode_fn() != nullptr; }   // the switch is on and the driver exports cuTensorMapEncodeTiled
int set_enabled(int on) {
  const int prev = enabled() ? 1 : 0;
  g_enabled.store(on ? 1 : 0, std::memory_order_relaxed);
  return prev;
}

// QG_OK: launched.  kNotTaken: the launch does not qualify (the caller runs the FMA kernel).
int launch(const Launch<float>& L, cudaStream_t st) {
  if (!enabled() || L.ntasks != 1) return kNotTaken;
  const Task<float>& t = L.t[0];
  if (t.skip_tm >= 0 || t.skip_tn >= 0 || t.K % KS || t.K < 64 || t.M < 128 || t.N < 96 || (t.N & 1) || L.batch > 65535)
    return kNotTaken;
  if ((t.ldc1 & 1) || (t.C2 && (t.ldc2 & 1)) || (t.X0 && (t.ldx0 & 1)) || (t.X1 && (t.ldx1 & 1)) || (t.X2 && (t.ldx2 & 1)))
    return kNotTaken;
  if (((uintptr_t)t.C1 & 15) || ((uintptr_t)t.C2 & 15) || ((uintptr_t)t.X0 & 15) || ((uintptr_t)t.X1 & 15) || ((uintptr_t)t.X2 & 15) ||
      (t.sC1 & 1) || (t.sC2 & 1) || (t.sX0 & 1) || (t.sX1 & 1) || (t.sX2 & 1))
    return kNotTaken;
  if (t.pred_it >= 0 && !L.s) return kNotTaken;
  CUtensorMap mA0, mB0, mA1, mB1;
  if (!make_map(&mA0, t.A0, t.K, t.M, L.batch, t.lda0, t.sA0, 2 * KS, TM, CU_TENSOR_MAP_SWIZZLE_128B)) return kNotTaken;
  if (!make_map(&mB0, t.B0, t.N, t.K, L.batch, t.ldb0, t.sB0, 2 * TN, KS, CU_TENSOR_MAP_SWIZZLE_NONE)) return kNotTaken;
  if (t.A1) {
    if (!make_map(&mA1, t.A1, t.K, t.M, L.batch, t.lda1, t.sA1, 2 * KS, TM, CU_TENSOR_MAP_SWIZZLE_128B)) return kNotTaken;
    if (!make_map(&mB1, t.B1, t.N, t.K, L.batch, t.ldb1, t.sB1, 2 * TN, KS, CU_TENSOR_MAP_SWIZZLE_NONE)) return kNotTaken;
  } else {
    mA1 = mA0; mB1 = mB0;
  }
  static std::atomic<unsigned long long> attr_done{0};
  if (first_use_on_device(attr_done)) {
    QG_RETURN_IF_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    QG_RETURN_IF_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  }
  Params P;
  P.t = t; P.batch = L.batch; P.s = L.s; P.sym = (t.herm && t.M == t.N) ? L.sym : nullptr; P.kstages = t.K / KS; P.nterms = t.A1 ? 2 : 1;
  static const int chunk = [] { const char* e = getenv("QG_TC_CHUNK"); const int v = e ? atoi(e) : 1; return v == 2 ? 2 : 1; }();   // tuning aid
  static const int no_debias = [] { const char* e = getenv("QG_TC_DEBIAS"); return (e && *e == '0') ? 1 : 0; }();                  // tuning aid
  P.debias = no_debias ? 0.f : (chunk == 1 ? 8.64e-8f : chunk == 2 ? 1.524e-7f : 0.f);
  dim3 grid((t.N + TN - 1) / TN, (t.M + TM - 1) / TM, (unsigned)L.batch);
  if (chunk == 2) gemm_tc_kernel<2><<<grid, NTHREADS, SMEM_BYTES, st>>>(mA0, mB0, mA1, mB1, P);
  else gemm_tc_kernel<1><<<grid, NTHREADS, SMEM_BYTES, st>>>(mA0, mB0, mA1, mB1, P);
  QG_CHECK_LAUNCH();
  return QG_OK;
}

}  // namespace tc
}  // namespace gemm

int gemm_tc_set_enabled(int on) { return gemm::tc::set_enabled(on); }

}  // namespace qg
